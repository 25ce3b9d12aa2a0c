timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench57.err | tee gpurun_out/bench_57.json | cut -c1-300
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench57.err | tee gpurun_out/bench_57_ref.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches57.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b57.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_kernel -c 1 -s 2 -f -o gpurun_out/fill57 python tools/prof_run.py cfg2 - 3 > gpurun_out/ncu57.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_wide -c 1 -s 1 -f -o gpurun_out/wide57 python tools/prof_run.py cfg2 - 3 >> gpurun_out/ncu57.log 2>&1
ls -la gpurun_out
