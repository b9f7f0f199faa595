"""The reference arm: the UNMODIFIED reference brought next to the repo (baseline/_ref, git-ignored, travels with
the gpurun snapshot) plus the runners that time it.  Measurement infrastructure only -- nothing under
cs231_capsule_yolo_traffic_sign_detection_b200/ imports this package."""
