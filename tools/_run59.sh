timeout 600 python -m pytest tests/test_scaling.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/scaling_run.py cfg2 - 3 cpu 2>&1 | tail -6
timeout 300 python tools/scaling_run.py cfg3 - 2 2>&1 | tail -3
