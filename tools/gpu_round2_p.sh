#!/bin/bash
# table-driven code expansion + early sequence copy at N=1: e2e against the number of host threads
mkdir -p gpurun_out
for t in 1 2 3 4 8; do ABEA_HOST_THREADS=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/rp_bench_t$t.json
  python -c "
import json
d=json.load(open('gpurun_out/rp_bench_t$t.json')); print('threads $t dev ms %.3f e2e ms %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d['e2e']['last_step_parts_ms_rank0'], d['e2e']['last_step_output_equals_resident_result_all_ranks'])"; done
ABEA_STREAM=3 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('STREAM=3 e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"
for t in 16 8 4; do echo "== dropin threads $t"; timeout 300 python tools/dropin_run.py cfg5 $t 6 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q 2>&1 | tail -2
