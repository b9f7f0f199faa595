"""B200-native capsule dynamic routing (the hot path of
Cranial-XIX/cs231-capsule-yolo-traffic-sign-detection), behind the reference's module interface.

    from cs231_capsule_yolo_traffic_sign_detection_b200 import CapsuleLayer
    import models; models.CapsuleLayer = CapsuleLayer      # then build CapsuleNet / DarkCapsuleNet

The arithmetic lives in libcaps_routing.so (hand-written sm_100a CUDA, C ABI in
include/caps_routing.h).  Importing the package does not need a GPU; running the routing branch
does, and fails loudly without the built library -- there is no CPU fallback."""
from . import _cabi
from .capsule import (CapsuleLayer, GraphedStep, HostPipe, HostStep, dark_capsule_loss, dark_regroup, dynamic_routing, primary_capsules, routing_margin_loss)
from .parallel import GradBucket, init_from_env, shard_bounds
from . import runner

__all__ = ['CapsuleLayer', 'GraphedStep', 'HostPipe', 'HostStep', 'dark_capsule_loss', 'dark_regroup', 'dynamic_routing', 'primary_capsules', 'routing_margin_loss',
           'GradBucket', 'init_from_env', 'shard_bounds', 'runner', '_cabi']
