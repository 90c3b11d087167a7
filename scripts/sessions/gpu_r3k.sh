#!/bin/bash
# GPU session r3k (2 GPUs): data-parallel D-MPNN engine against a single-GPU replica; GraphConv bench at N=2; gloo-free
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/dmpnn_dp2.py > gpurun_out/r3k_dmpnn_dp2.json 2> gpurun_out/r3k_dmpnn_dp2.err; echo "dmpnn dp2 exit $?"; cat gpurun_out/r3k_dmpnn_dp2.json; tail -n 4 gpurun_out/r3k_dmpnn_dp2.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3k_bench_n2.json 2> gpurun_out/r3k_bench_n2.err; echo "bench n2 exit $?"
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r3k_bench_n2.json") if l.startswith("{")][-1])
print("n=%d value %.0f ms %.4f e2e %.0f (%.4f ms)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
P
