#!/bin/bash
# thresholds anchored to the measured cycle counts: the e2e leg must schedule like the resident one (6 wide reads)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/rt_bench.json
python -c "
import json
d=json.load(open('gpurun_out/rt_bench.json')); print('dev ms %.3f e2e ms %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), 'wide', d['reads_wide_rank0'], d['scheduler_model_cycles'], d['e2e']['last_step_parts_ms_rank0'], d['e2e']['last_step_output_equals_resident_result_all_ranks'])"
for c in cfg3 cfg4 cfg2; do timeout 300 python tools/prof_run.py $c - 3 2>&1 | grep -E "kernel_ms" | tail -1; done
timeout 300 python tools/dropin_run.py cfg5 16 6 2>&1 | tail -1
