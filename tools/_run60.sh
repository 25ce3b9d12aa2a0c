timeout 300 python tools/scaling_run.py cfg2 - 3 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_scaling_kernel -c 1 -f -o gpurun_out/scl60 python tools/scaling_run.py cfg2 - 1 > gpurun_out/ncu60.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_mom_kernel -c 1 -f -o gpurun_out/mom60 python tools/scaling_run.py cfg2 - 1 >> gpurun_out/ncu60.log 2>&1
