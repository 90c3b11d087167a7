#!/bin/bash
mkdir -p gpurun_out
for ko in 29 21 13; do
echo "== knockout $ko"
DCGC_TC_KNOCKOUT=$ko timeout 120 python scripts/gemm_timeline.py 2>&1 | head -n 13 | cut -c1-500 | grep -E "tf32x3|mma chunk|^ +[0-5] \|"
done
