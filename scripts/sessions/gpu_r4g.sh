#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/r4g_pytest.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/r4g_pytest.log | tail -n 3; grep -E "^FAILED|Error|^E  " gpurun_out/r4g_pytest.log | cut -c1-250 | head -n 10
timeout 120 python scripts/gemm_timeline.py > gpurun_out/r4g_gemm_timeline_v4.log 2>&1; echo "timeline exit $?"; head -n 14 gpurun_out/r4g_gemm_timeline_v4.log | cut -c1-700
for ko in 0 1 4; do
DCGC_TC_KNOCKOUT=$ko timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-e2e --sub "" --breakdown gpurun_out/r4g_breakdown_ko$ko.md > gpurun_out/r4g_bench_ko$ko.json 2> gpurun_out/r4g_ko$ko.err; echo "knockout $ko exit $?"; grep -E "timed|gemm_fwd|gemm_dgrad|linear_fwd" gpurun_out/r4g_breakdown_ko$ko.md
done
