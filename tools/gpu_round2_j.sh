#!/bin/bash
# wide-kernel hand-over A/B: mbarrier (mbar) vs tagged publish with 4 (default lib) or 2 (tag2) steps per trip
mkdir -p gpurun_out
for c in cfg3 cfg5; do
  for v in default mbar tag2; do
    if [ $v = default ]; then unset ABEA_LIB; else export ABEA_LIB=$PWD/f5c_b200/lib/exp/libabea_$v.so; fi
    timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/rj_prof_${c}_$v.txt 2>&1
    echo "== $c $v"; grep -E "kernel_ms|fill cycles" gpurun_out/rj_prof_${c}_$v.txt | tail -3
  done
done
unset ABEA_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
