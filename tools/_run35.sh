for s in 1 2; do ABEA_STREAM=$s python tools/e2e_run.py cfg2 - 5; done
for lc in 148 296 592; do ABEA_STREAM=1 ABEA_LOAD_CTAS=$lc python tools/e2e_run.py cfg2 - 5; done
ABEA_STREAM=3 ABEA_LOAD_CTAS=296 python tools/e2e_run.py cfg2 - 5
