timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export E2E_STARTS=1
for s in 0 1 2 3; do ABEA_STREAM=$s timeout 120 python tools/e2e_run.py cfg2 - 5 | head -1; done
for lc in 32 148; do ABEA_STREAM=3 ABEA_LOAD_CTAS=$lc timeout 120 python tools/e2e_run.py cfg2 - 5 | head -1; done
timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench_40.json | cut -c1-400
