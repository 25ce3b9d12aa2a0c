#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r7_tests.log
ABEA_TIME_PACK=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
ABEA_TIME_PACK=1 timeout 300 python tools/dropin_run.py cfg5 16 5 > gpurun_out/r7_dropin_t16.txt 2>&1
for c in cfg5 cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/r7_prof_$c.txt 2>&1; done
ABEA_LIB=$PWD/f5c_b200/lib/libabea_b200_w16.so ABEA_FILL_WARPS_PER_CTA=16 ABEA_CARVEOUT=45 ABEA_WIDE_CAP=6 timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r7_prof_cfg5_w16.txt 2>&1
tail -4 gpurun_out/r7_tests.log; python -c "
import json
d=json.load(open('gpurun_out/r7_bench.json')); print('dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0']); print('dropin', d.get('e2e_dropin',{}).get('ms_per_step'), 'gpuref', d.get('gpu_reference',{}).get('ms_per_step'), 'parity', d.get('parity_on_cpu_sample')); print([(c['config'], round(c['ms_per_step'],2), c['parity_on_cpu_sample']) for c in d.get('configs',[])])"
tail -3 gpurun_out/r7_dropin_t16.txt; tail -8 gpurun_out/r7_prof_cfg5.txt; grep -E "kernel_ms" gpurun_out/r7_prof_cfg3.txt gpurun_out/r7_prof_cfg4.txt | tail -2; grep -E "kernel_ms|median" gpurun_out/r7_prof_cfg5_w16.txt | tail -3
