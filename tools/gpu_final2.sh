#!/bin/bash
# the GPU tests and smoke on the final tree + memcheck over every kernel incl. the ex-zd decoder
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests_final2.log; cat gpurun_out/tests_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck_final2.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck_final2.log
