#!/bin/bash
# GPU session r1k: D-MPNN path first run + BN/pool micro-optimisations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dmpnn.py -m gpu -q -x > gpurun_out/r1k_pytest_dmpnn.log 2>&1; echo "pytest dmpnn exit $?"
tail -40 gpurun_out/r1k_pytest_dmpnn.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dmpnn.py > gpurun_out/r1k_pytest.log 2>&1; echo "pytest rest exit $?"
tail -5 gpurun_out/r1k_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline --breakdown gpurun_out/r1k_breakdown_tc.md > gpurun_out/r1k_bench_tc.json 2> gpurun_out/r1k_bench_tc.err; echo "bench tc exit $?"
cut -c1-400 gpurun_out/r1k_bench_tc.json; tail -3 gpurun_out/r1k_bench_tc.err; cat gpurun_out/r1k_breakdown_tc.md
