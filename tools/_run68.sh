for r in 0 1 2 4; do echo "ABEA_SM_RESERVE=$r (2-GPU shard)"; ABEA_SM_RESERVE=$r PROF_WORLD=2 timeout 300 python tools/wide_var.py cfg2 4 2>&1 | tail -4 | cut -c1-150; done
for r in 0 1 2; do echo "ABEA_SM_RESERVE=$r (cfg2)"; ABEA_SM_RESERVE=$r timeout 300 python tools/wide_var.py cfg2 4 2>&1 | tail -4 | cut -c1-150; done
