#!/bin/bash
# GPU session r3b: full GPU suite (incl. real-molecule Delaney tests, sliced EdgeNetwork contraction) + MPNN bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/r3b_pytest.log
timeout 300 python scripts/bench_mpnn.py > gpurun_out/r3b_mpnn.log 2>&1; echo "mpnn exit $?"; tail -n 5 gpurun_out/r3b_mpnn.log
