for kb in 96 48 24 16 8; do for lc in 64 128; do ABEA_LOAD_PIECE_KB=$kb ABEA_LOAD_CTAS=$lc timeout 120 python tools/e2e_run.py cfg2 - 5 | head -1 | sed "s/^/piece_kb=$kb /"; done; done
for crit in 0.8 0.6; do ABEA_LOAD_PIECE_KB=16 ABEA_LOAD_CRIT=$crit timeout 120 python tools/e2e_run.py cfg2 - 5 | head -1 | sed "s/^/piece_kb=16 crit=$crit /"; done
