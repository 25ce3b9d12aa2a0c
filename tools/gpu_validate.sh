#!/bin/bash
# The round's standard validation pass on a GPU box (run under gpurun from the repo root): parity tests, smoke, both
# bench arms, the ncu launch list of the bench command, ncu --set full captures of the dominant (narrow) kernel and of
# the wide kernel, compute-sanitizer memcheck / racecheck over every kernel. Outputs land in gpurun_out/ (scratch);
# summaries are made here afterwards (tools/ncu_summary.py, opcode_mix.py, traffic_json.py) and copied to profiles/.
#   gpurun --timeout 2700 -- 'bash tools/gpu_validate.sh TAG'
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_${TAG}_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/b_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_kernel -c 1 -s 2 -f -o gpurun_out/fill_$TAG python tools/prof_run.py cfg5 - 3 > gpurun_out/ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_wide -c 1 -s 2 -f -o gpurun_out/wide_$TAG python tools/prof_run.py cfg5 - 3 >> gpurun_out/ncu_$TAG.log 2>&1
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_${tool}_$TAG.log; done
for m in 3 4; do ABEA_STREAM=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_${TAG}_stream$m.json; done
for c in cfg5 cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/prof_${TAG}_$c.txt 2>&1; done
cat gpurun_out/tests_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; cut -c1-400 gpurun_out/bench_$TAG.json; cut -c1-200 gpurun_out/bench_${TAG}_ref.json
for m in 3 4; do python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}_stream$m.json')); print('ABEA_STREAM=$m e2e ms %.3f'%d['e2e']['ms_per_step'], d['e2e']['last_step_parts_ms_rank0'])"; done
