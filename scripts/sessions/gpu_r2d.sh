#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1 2 3; do for w in gather pool_fwd; do timeout 120 python scripts/mg_timeline.py $w $mode > gpurun_out/r2d_timeline_${w}_$mode.txt 2>&1; echo "mode $mode: $(head -2 gpurun_out/r2d_timeline_${w}_$mode.txt | tr '\n' ' ')"; done; done
for mode in 1 2 3; do echo "== mode $mode"; sed -n 3,12p gpurun_out/r2d_timeline_gather_$mode.txt; done
