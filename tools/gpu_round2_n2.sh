#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/rn_bench_n$N.json 2> gpurun_out/rn_bench_n$N.err
tail -5 gpurun_out/rn_bench_n$N.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/rn_bench_n$N.json')); print('N=$N value %.3e'%d['value'], 'dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'gather ms %.3f'%d['e2e']['nccl_gather_ms_per_step'], 'load_ms max %.2f'%d['e2e']['load_ms_max_over_ranks'], d.get('parity_gathered_shard'), d['e2e']['last_step_output_equals_resident_result_all_ranks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --cpu-sample-events 2e6 2>/dev/null | cut -c1-250
