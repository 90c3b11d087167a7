#!/bin/bash
# GPU session r1i: fused BN statistics without epilogue unrolling
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1i_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r1i_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline --breakdown gpurun_out/r1i_breakdown_tc.md > gpurun_out/r1i_bench_tc.json 2> gpurun_out/r1i_bench_tc.err; echo "bench tc exit $?"
cat gpurun_out/r1i_bench_tc.json; tail -3 gpurun_out/r1i_bench_tc.err; cat gpurun_out/r1i_breakdown_tc.md
