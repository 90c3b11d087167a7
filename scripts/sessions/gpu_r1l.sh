#!/bin/bash
# GPU session r1l (2 GPUs): tests, 1-GPU bench, 2-GPU data-parallel bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1l_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r1l_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 --gemm-mode tf32x3 > gpurun_out/r1l_bench_n1.json 2> gpurun_out/r1l_bench_n1.err; echo "bench n1 exit $?"
cut -c1-1500 gpurun_out/r1l_bench_n1.json; tail -3 gpurun_out/r1l_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline > gpurun_out/r1l_bench_n2.json 2> gpurun_out/r1l_bench_n2.err; echo "bench n2 exit $?"
cut -c1-1500 gpurun_out/r1l_bench_n2.json; tail -5 gpurun_out/r1l_bench_n2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1l_bench_ref.json 2> gpurun_out/r1l_bench_ref.err; echo "bench ref exit $?"
cut -c1-600 gpurun_out/r1l_bench_ref.json
