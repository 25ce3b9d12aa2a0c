#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" ABEA_TIME_PACK=1 timeout 300 python tools/dropin_run.py cfg5 ${THREADS:-16} 5 > gpurun_out/r8_dropin_$name.txt 2>&1; echo "== $name"; grep "ragged" gpurun_out/r8_dropin_$name.txt | tail -2; tail -1 gpurun_out/r8_dropin_$name.txt; }
run base X=1
THREADS=4 run t4 X=1
THREADS=32 run t32 X=1
run latecopy ABEA_RAG_LATE_COPY=1
run poll50 ABEA_RAG_POLL_US=50
run load16 ABEA_LOAD_CTAS=16
run carve100 ABEA_CARVEOUT=100
run piece32 ABEA_LOAD_PIECE_KB=32
run nowide ABEA_WIDE=0
ABEA_LIB=$PWD/f5c_b200/lib/libabea_b200_widef32.so timeout 300 python tools/prof_run.py cfg3 - 3 > gpurun_out/r8_prof_cfg3_widef32.txt 2>&1; grep -E "kernel_ms|fill cycles" gpurun_out/r8_prof_cfg3_widef32.txt | tail -3
