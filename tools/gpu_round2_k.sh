#!/bin/bash
# scheduler thresholds re-swept with the round-2 kernels (narrow 856 / 625 alone, wide 400 cycles per band)
mkdir -p gpurun_out
export ABEA_LIB=$PWD/f5c_b200/lib/exp/libabea_mbar.so
timeout 900 python tools/sweep_run.py cfg5 ";ABEA_WIDE_ALPHA=0.7;ABEA_WIDE_ALPHA=0.6;ABEA_WIDE_ALPHA=0.5;ABEA_WIDE_ALPHA=0.4;ABEA_LONG_ALPHA=0.6;ABEA_LONG_ALPHA=1.0;ABEA_WIDE_ALPHA=0.6 ABEA_LONG_ALPHA=0.6;ABEA_WIDE_ALPHA=0.5 ABEA_LONG_ALPHA=0.6;ABEA_WIDE_ALPHA=0.6 ABEA_LONG_ALPHA=1.0;ABEA_WIDE_ALPHA=0.5 ABEA_LONG_ALPHA=1.2;ABEA_WIDE_ALPHA=0.5 ABEA_LONG_ALPHA=2.0" 4 > gpurun_out/rk_sweep_cfg5.txt 2>&1
cat gpurun_out/rk_sweep_cfg5.txt
timeout 600 python tools/sweep_run.py cfg2 ";ABEA_WIDE_ALPHA=0.6;ABEA_WIDE_ALPHA=0.5;ABEA_WIDE_ALPHA=0.5 ABEA_LONG_ALPHA=1.2" 4 > gpurun_out/rk_sweep_cfg2.txt 2>&1
cat gpurun_out/rk_sweep_cfg2.txt
