timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none -k regex:abea_events -c 6 --csv --log-file gpurun_out/evt66.csv python tools/events_run.py 4096 4000 2 > gpurun_out/ncu66.log 2>&1
grep -v "^==" gpurun_out/evt66.csv | cut -d, -f5,13- | head -30
