#!/bin/bash
# where the streamed step loses its 0.75 ms against the resident one: per-read start / end times of both
mkdir -p gpurun_out
E2E_STARTS=1 timeout 300 python tools/e2e_run.py cfg5 - 6 > gpurun_out/rs_streamed.txt 2>&1; cat gpurun_out/rs_streamed.txt
E2E_STARTS=1 E2E_RESIDENT=1 timeout 300 python tools/e2e_run.py cfg5 - 2 > gpurun_out/rs_resident.txt 2>&1; cat gpurun_out/rs_resident.txt
