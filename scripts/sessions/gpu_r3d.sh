#!/bin/bash
# GPU session r3d: fused D-MPNN engine: tests, then throughput with / without the engine
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dmpnn.py -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest exit $?"
tail -n 25 gpurun_out/r3d_pytest.log | cut -c1-300
timeout 300 python scripts/bench_extra.py dmpnn > gpurun_out/r3d_dmpnn_engine.json 2> gpurun_out/r3d_dmpnn_engine.err; echo "engine exit $?"
cat gpurun_out/r3d_dmpnn_engine.json; tail -n 3 gpurun_out/r3d_dmpnn_engine.err
DCGC_DMPNN_ENGINE=0 timeout 300 python scripts/bench_extra.py dmpnn > gpurun_out/r3d_dmpnn_autograd.json 2> gpurun_out/r3d_dmpnn_autograd.err; echo "autograd exit $?"
cat gpurun_out/r3d_dmpnn_autograd.json
