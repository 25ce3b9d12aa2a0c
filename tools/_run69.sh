echo "2-GPU shard"; PROF_WORLD=2 timeout 300 python tools/wide_var.py cfg2 8 2>&1 | tail -8 | cut -c1-150
echo "cfg2"; timeout 300 python tools/wide_var.py cfg2 8 2>&1 | tail -8 | cut -c1-150
for c in cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 2 2>&1 | head -2; done
