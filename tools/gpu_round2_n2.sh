#!/bin/bash
# two ranks on one node as the driver launches them (our arm only): the NCCL result exchange on hardware
mkdir -p gpurun_out
N=${1:-2}
nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/rq_bench_n$N.json 2> gpurun_out/rq_bench_n$N.err
tail -5 gpurun_out/rq_bench_n$N.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/rq_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value %.3e'%d['value'], 'dev ms %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'gather ms %.3f'%d['e2e']['nccl_gather_ms_per_step'], 'load_ms max %.2f'%d['e2e']['load_ms_max_over_ranks'], 'threads', d['e2e'].get('host_threads'), d.get('parity_gathered_shard'), d['e2e']['last_step_output_equals_resident_result_all_ranks'], d['e2e']['last_step_parts_ms_rank0'])"
