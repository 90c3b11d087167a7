#!/bin/bash
# 2-GPU e2e study: effect of the number of host layout workers per rank
mkdir -p gpurun_out
for W in 1 2 4; do
DCGC_HOST_WORKERS=$W timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$W bench.py --gpus 2 --steps 30 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('workers=$W n=2 value %.0f ms %.3f e2e %.0f e2e_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
"
done
DCGC_HOST_WORKERS=2 timeout 600 python bench.py --steps 30 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('workers=2 n=1 value %.0f ms %.3f e2e %.0f e2e_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
"
