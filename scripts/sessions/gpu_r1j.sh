#!/bin/bash
# GPU session r1j: ncu full capture of the HBM-bound kernels inside one engine step
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --gemm-mode tf32x3"
timeout 300 $CMD > gpurun_out/r1j_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"col_moments|pool_fwd|pool_bwd|bn_relu|gather_fwd|gather_bwd" -s 22 -c 11 -o gpurun_out/r1j_hbm $CMD > gpurun_out/r1j_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r1j_ncu.log
