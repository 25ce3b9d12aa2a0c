#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r4_tests.log
ABEA_TIME_PACK=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
for cv in 60 35; do ABEA_CARVEOUT=$cv timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r4_bench_carve$cv.json 2> gpurun_out/r4_bench_carve$cv.err; done
for c in cfg5 cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/r4_prof_$c.txt 2>&1; done
ABEA_CARVEOUT=35 timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r4_prof_cfg5_carve35.txt 2>&1
timeout 300 python tools/blow5_run.py reads.blow5 16 > gpurun_out/r4_blow5.txt 2>&1
timeout 300 python tools/blow5_run.py ecoli8_zlib_svbzd.blow5 128 >> gpurun_out/r4_blow5.txt 2>&1
tail -4 gpurun_out/r4_tests.log; for f in gpurun_out/r4_bench.json gpurun_out/r4_bench_carve60.json gpurun_out/r4_bench_carve35.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', 'dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"; done
grep "abea pack" gpurun_out/r4_bench.err | tail -3; tail -8 gpurun_out/r4_prof_cfg5.txt; grep -E "kernel_ms|trace cycles" gpurun_out/r4_prof_cfg3.txt gpurun_out/r4_prof_cfg4.txt gpurun_out/r4_prof_cfg5_carve35.txt | tail -12; cat gpurun_out/r4_blow5.txt
