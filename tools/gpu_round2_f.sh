#!/bin/bash
mkdir -p gpurun_out
for t in 16 8; do ABEA_TIME_PACK=1 timeout 300 python tools/dropin_run.py cfg5 $t 5 > gpurun_out/r6_dropin_t$t.txt 2>&1; done
ABEA_CARVEOUT=100 ABEA_TIME_PACK=1 timeout 300 python tools/dropin_run.py cfg5 16 5 > gpurun_out/r6_dropin_carve100.txt 2>&1
ABEA_LIB=$PWD/f5c_b200/lib/libabea_b200_w16.so ABEA_FILL_WARPS_PER_CTA=16 ABEA_CARVEOUT=45 timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r6_prof_cfg5_w16.txt 2>&1
ABEA_LIB=$PWD/f5c_b200/lib/libabea_b200_w16.so ABEA_FILL_WARPS_PER_CTA=16 ABEA_CARVEOUT=45 ABEA_LONG_ALPHA=100 timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r6_prof_cfg5_w16_nopause.txt 2>&1
ABEA_LONG_ALPHA=100 timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r6_prof_cfg5_nopause.txt 2>&1
ABEA_FILL_WARPS_PER_CTA=8 timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r6_prof_cfg5_w8.txt 2>&1
for f in gpurun_out/r6_dropin_t16.txt gpurun_out/r6_dropin_t8.txt gpurun_out/r6_dropin_carve100.txt; do echo == $f; tail -4 $f; done
for f in gpurun_out/r6_prof_cfg5_w16.txt gpurun_out/r6_prof_cfg5_w16_nopause.txt gpurun_out/r6_prof_cfg5_nopause.txt gpurun_out/r6_prof_cfg5_w8.txt; do echo == $f; grep -E "kernel_ms|median reads" $f | tail -3; done
