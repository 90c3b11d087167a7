#!/bin/bash
# GPU session r1d: tcgen05 wgrad + streamed-weight GEMM v3 + new BN statistics kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/r1d_pytest_tc.log 2>&1; echo "pytest tc exit $?"
tail -25 gpurun_out/r1d_pytest_tc.log
if grep -q failed gpurun_out/r1d_pytest_tc.log; then
  DCGC_TC_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/r1d_pytest_tc_v2.log 2>&1; echo "pytest tc (variant 2) exit $?"
  tail -15 gpurun_out/r1d_pytest_tc_v2.log
fi
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tc.py > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rest exit $?"
tail -8 gpurun_out/r1d_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline --breakdown gpurun_out/r1d_breakdown_tc.md > gpurun_out/r1d_bench_tc.json 2> gpurun_out/r1d_bench_tc.err; echo "bench tc exit $?"
cat gpurun_out/r1d_bench_tc.json; tail -3 gpurun_out/r1d_bench_tc.err; cat gpurun_out/r1d_breakdown_tc.md
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --breakdown gpurun_out/r1d_breakdown_fp32.md > gpurun_out/r1d_bench_fp32.json 2> gpurun_out/r1d_bench_fp32.err; echo "bench fp32 exit $?"
cat gpurun_out/r1d_bench_fp32.json; cat gpurun_out/r1d_breakdown_fp32.md
