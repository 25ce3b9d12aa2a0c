#!/bin/bash
# path codes out (ABEA_STREAM bit 2): parity, then the e2e legs against the whole-list form (ABEA_STREAM=3)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/rm_tests.log; cat gpurun_out/rm_tests.log
for m in 5 3 4; do
  ABEA_STREAM=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/rm_bench_s$m.json
  python -c "
import json
d=json.load(open('gpurun_out/rm_bench_s$m.json')); print('ABEA_STREAM=$m dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'], 'd2h', d['e2e']['d2h_bytes_per_step'], 'ok', d['e2e']['last_step_output_equals_resident_result_all_ranks'])"
done
for t in 4 8 16; do ABEA_HOST_THREADS=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/rm_bench_t$t.json
  python -c "
import json
d=json.load(open('gpurun_out/rm_bench_t$t.json')); print('threads $t e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"; done
for s in 5 3; do for t in 16 8; do echo "== dropin STREAM=$s threads $t"; ABEA_STREAM=$s ABEA_TIME_PACK=1 timeout 300 python tools/dropin_run.py cfg5 $t 6 > gpurun_out/rm_dropin_s${s}_t$t.txt 2>&1; grep ragged gpurun_out/rm_dropin_s${s}_t$t.txt | tail -1; tail -1 gpurun_out/rm_dropin_s${s}_t$t.txt; done; done
