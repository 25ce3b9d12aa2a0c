timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_events_kernel -c 1 -f -o gpurun_out/evt64 python tools/events_run.py 4096 4000 1 > gpurun_out/ncu64.log 2>&1
tail -3 gpurun_out/ncu64.log
