#!/bin/bash
# 2-GPU session: tests, then N=1 and N=2 bench (value + e2e) with the slot-based prefetch pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1n_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r1n_pytest.log
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("n=%d value %.0f ms %.3f e2e %.0f e2e_ms %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))'
for rep in 1 2; do
timeout 400 python bench.py --steps 40 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline 2>gpurun_out/r1n_n1.err | tee gpurun_out/r1n_bench_n1.json | python -c "$show"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$rep bench.py --gpus 2 --steps 40 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline 2>gpurun_out/r1n_n2.err | tee gpurun_out/r1n_bench_n2.json | python -c "$show"
done
tail -3 gpurun_out/r1n_n1.err
