export E2E_STARTS=1
ABEA_STREAM=2 python tools/e2e_run.py cfg2 - 5 | head -1
for lc in 32 64 148; do ABEA_STREAM=3 ABEA_LOAD_CTAS=$lc python tools/e2e_run.py cfg2 - 5; done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-250
