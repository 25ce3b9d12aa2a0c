#!/bin/bash
# round 2, run b: parallel traceback A/B, margin sweep, ncu of the dominant kernel on the target config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
for tb in 0 1; do ABEA_TB=$tb timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_tb$tb.json 2> gpurun_out/r2_bench_tb$tb.err; done
for c in cfg5 cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/r2_prof_$c.txt 2>&1; done
for m in 16 32 128; do ABEA_TB_MARGIN=$m timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r2_prof_cfg5_m$m.txt 2>&1; done
ABEA_TB_MARGIN=128 timeout 300 python tools/prof_run.py cfg3 - 3 > gpurun_out/r2_prof_cfg3_m128.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_kernel -c 1 -s 2 -f -o gpurun_out/r2_fill_cfg5 python tools/prof_run.py cfg5 - 3 > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_tests.log; for tb in 0 1; do cut -c1-330 gpurun_out/r2_bench_tb$tb.json; done; tail -8 gpurun_out/r2_prof_cfg5.txt; tail -8 gpurun_out/r2_prof_cfg3.txt
