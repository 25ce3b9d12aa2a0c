timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench67.err | tee gpurun_out/bench_67.json | cut -c1-200
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench67.err | tee gpurun_out/bench_67_ref.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches67.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b67.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:abea_fill_wide -c 1 -s 2 -f -o gpurun_out/wide67 python tools/prof_run.py cfg2 - 3 > gpurun_out/ncu67.log 2>&1
tail -2 gpurun_out/bench67.err
