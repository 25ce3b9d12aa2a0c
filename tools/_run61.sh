timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/scaling_run.py cfg2 - 3 cpu 2>&1 | tail -6
timeout 300 python tools/scaling_run.py cfg3 - 2 2>&1 | tail -3
timeout 300 python tools/scaling_run.py cfg4 - 2 2>&1 | tail -3
