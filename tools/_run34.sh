python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in 0 1; do ABEA_STREAM=$s python tools/e2e_run.py cfg2 - 6; done
for lc in 16 32 128 148; do ABEA_LOAD_CTAS=$lc python tools/e2e_run.py cfg2 - 6; done
for c in cfg3 cfg4; do for s in 0 1; do ABEA_STREAM=$s python tools/e2e_run.py $c - 3; done; done
