#!/bin/bash
# what the pair lists cost on their way out: the copy to the caller's mapped buffer cut to 1/2 and 1/64 of every list
# (experimental builds; results of those runs are not valid alignments), in the streamed (3) and copy-engine-in (2) modes
mkdir -p gpurun_out
for v in default fin2 fin64; do
  if [ $v = default ]; then unset ABEA_LIB; else export ABEA_LIB=$PWD/f5c_b200/lib/exp/libabea_$v.so; fi
  for m in 3 2; do
    echo "== $v STREAM=$m"; ABEA_STREAM=$m timeout 300 python tools/e2e_run.py cfg5 - 8 2>&1 | tail -1
  done
done > gpurun_out/rl_fin.txt 2>&1
cat gpurun_out/rl_fin.txt
