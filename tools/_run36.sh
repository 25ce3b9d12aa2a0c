export E2E_STARTS=1
ABEA_STREAM=2 python tools/e2e_run.py cfg2 - 3
ABEA_STREAM=3 ABEA_LOAD_CTAS=64 python tools/e2e_run.py cfg2 - 3
ABEA_STREAM=3 ABEA_LOAD_CTAS=148 python tools/e2e_run.py cfg2 - 3
