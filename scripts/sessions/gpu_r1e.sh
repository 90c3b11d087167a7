#!/bin/bash
# GPU session r1e: knockout study of the tcgen05 GEMM + ncu full capture of the TC kernels
mkdir -p gpurun_out
timeout 900 python scripts/gemm_knockout.py > gpurun_out/r1e_knockout.md 2> gpurun_out/r1e_knockout.err; echo "knockout exit $?"
cat gpurun_out/r1e_knockout.md; tail -5 gpurun_out/r1e_knockout.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --gemm-mode tf32x3"
timeout 300 $CMD > gpurun_out/r1e_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel_v3|tc_wgrad" -c 8 -o gpurun_out/r1e_tc $CMD > gpurun_out/r1e_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r1e_ncu.log
