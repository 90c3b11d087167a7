#!/bin/bash
# GPU session r3l (2 GPUs): GraphConv bench at N=2 after the lockstep fix; D-MPNN data-parallel diagnostics
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3l_bench_n2.json 2> gpurun_out/r3l_bench_n2.err; echo "bench n2 exit $?"
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r3l_bench_n2.json") if l.startswith("{")][-1])
    print("n=%d value %.0f ms %.4f e2e %.0f (%.4f ms)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("no bench line", e)
P
DP2_NO_TIMING=1 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/dmpnn_dp2.py > gpurun_out/r3l_dmpnn_dp2.json 2> gpurun_out/r3l_dmpnn_dp2.err; echo "dmpnn dp2 exit $?"; cut -c1-3000 gpurun_out/r3l_dmpnn_dp2.json; tail -n 5 gpurun_out/r3l_dmpnn_dp2.err | cut -c1-300
