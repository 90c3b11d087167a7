#!/bin/bash
# GPU session r1f: coalesced GEMM epilogue, row-block gather/pool kernels, vectorised gather_bwd
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/r1f_pytest.log
timeout 900 python scripts/gemm_knockout.py > gpurun_out/r1f_knockout.md 2> gpurun_out/r1f_knockout.err; echo "knockout exit $?"
cat gpurun_out/r1f_knockout.md; tail -5 gpurun_out/r1f_knockout.err
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline --breakdown gpurun_out/r1f_breakdown_tc.md > gpurun_out/r1f_bench_tc.json 2> gpurun_out/r1f_bench_tc.err; echo "bench tc exit $?"
cat gpurun_out/r1f_bench_tc.json; tail -3 gpurun_out/r1f_bench_tc.err; cat gpurun_out/r1f_breakdown_tc.md
