timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/sweep.py --quick 2>&1 | tail -5
