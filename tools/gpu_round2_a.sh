#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r1_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
for c in cfg5 cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/r1_prof_$c.txt 2>&1; done
tail -5 gpurun_out/r1_tests.log; tail -3 gpurun_out/r1_smoke.log; cut -c1-600 gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
