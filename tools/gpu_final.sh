#!/bin/bash
# last pass of the round (the GPU budget does not cover the whole of gpu_validate.sh again): bench both arms, launch list,
# one ncu --set full capture of the dominant kernel, memcheck over every kernel incl. the code kernels
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_${TAG}_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/b_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:abea_fill_kernel -c 1 -s 2 -f -o gpurun_out/fill_$TAG python tools/prof_run.py cfg5 - 3 > gpurun_out/ncu_$TAG.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
cut -c1-300 gpurun_out/bench_$TAG.json; cut -c1-200 gpurun_out/bench_${TAG}_ref.json
