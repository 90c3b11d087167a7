#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1 2 3; do for w in gather pool_fwd; do timeout 120 python scripts/mg_timeline.py $w $mode > gpurun_out/r2a_timeline_${w}_$mode.txt 2>&1; echo "mode $mode: $(head -1 gpurun_out/r2a_timeline_${w}_$mode.txt)"; done; done
for mode in 1 2 3; do echo "== mode $mode"; head -12 gpurun_out/r2a_timeline_gather_$mode.txt; done
