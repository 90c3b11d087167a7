#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/gemm_timeline.py > gpurun_out/r4f_gemm_timeline_v4.log 2>&1; echo "timeline exit $?"; head -n 14 gpurun_out/r4f_gemm_timeline_v4.log | cut -c1-400
for ko in 0 1 4 5; do
DCGC_TC_KNOCKOUT=$ko timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-e2e --sub "" --breakdown gpurun_out/r4f_breakdown_ko$ko.md > /dev/null 2> gpurun_out/r4f_ko$ko.err; echo "knockout $ko exit $?"; grep -E "gemm_fwd|gemm_dgrad|linear_fwd" gpurun_out/r4f_breakdown_ko$ko.md
done
