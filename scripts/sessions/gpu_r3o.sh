#!/bin/bash
# GPU session r3o (1 GPU): GPU suite at HEAD; where a small-batch step goes (host stage timers + device scopes)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r3o_pytest.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/r3o_pytest.log | cut -c1-300
timeout 100 python scripts/small_batch_trace.py > gpurun_out/r3o_small_batch.json 2> gpurun_out/r3o_small_batch.err; echo "trace exit $?"; cut -c1-1200 gpurun_out/r3o_small_batch.json; tail -n 3 gpurun_out/r3o_small_batch.err | cut -c1-300
