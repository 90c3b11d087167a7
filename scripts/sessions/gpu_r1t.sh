#!/bin/bash
# 8-GPU weak-scaling check of the default bench command
mkdir -p gpurun_out
nproc; nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r1t_bench_n8.json 2> gpurun_out/r1t_bench_n8.err; echo "bench n8 exit $?"
python -c "
import json
for l in open('gpurun_out/r1t_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n=%d value %.0f ms %.3f e2e %.0f e2e_ms %.3f workers %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('host_workers')))"
tail -3 gpurun_out/r1t_bench_n8.err
