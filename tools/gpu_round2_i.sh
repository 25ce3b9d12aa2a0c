#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r9_tests.log; tail -3 gpurun_out/r9_tests.log
ABEA_TIME_PACK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r9_bench.json')); print('dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"; grep "abea pack" gpurun_out/r9_bench.err | tail -2
run() { name=$1; shift; env "$@" ABEA_TIME_PACK=1 timeout 300 python tools/dropin_run.py cfg5 16 5 > gpurun_out/r9_dropin_$name.txt 2>&1; echo "== $name"; grep "ragged" gpurun_out/r9_dropin_$name.txt | tail -1; tail -1 gpurun_out/r9_dropin_$name.txt; }
run base X=1
run c148 ABEA_LOAD_CTAS=148
run c296 ABEA_LOAD_CTAS=296
run c296p16 ABEA_LOAD_CTAS=296 ABEA_LOAD_PIECE_KB=16
run c148p32 ABEA_LOAD_CTAS=148 ABEA_LOAD_PIECE_KB=32
run p64 ABEA_LOAD_PIECE_KB=64
for cta in 148 296; do ABEA_LOAD_CTAS=$cta timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r9_bench_c$cta.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r9_bench_c$cta.json')); print('load_ctas $cta: dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"; done
ABEA_LOAD_PIECE_KB=32 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r9_bench_p32.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r9_bench_p32.json')); print('piece 32K: dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"
