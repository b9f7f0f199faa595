timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_primary.py 2>&1 | tail -14 | tee gpurun_out/primary.md
