timeout 900 python -m pytest tests/test_events.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/events_run.py 4096 4000 3 cpu 2>&1 | tail -5
