#!/bin/bash
# The round's standard validation pass on a GPU box (run under gpurun from the repo root):
# parity tests, smoke, both bench arms, the ncu launch list of the bench command, and ncu --set full captures of the
# dominant kernel and of the wide kernel. Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
#   gpurun --timeout 1800 -- 'bash tools/gpu_validate.sh TAG'
TAG=${1:-run}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json | cut -c1-200
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_${TAG}_ref.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_kernel -c 1 -s 2 -f -o gpurun_out/fill_$TAG python tools/prof_run.py cfg2 - 3 > gpurun_out/ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_wide -c 1 -s 2 -f -o gpurun_out/wide_$TAG python tools/prof_run.py cfg2 - 3 >> gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/bench_$TAG.err
