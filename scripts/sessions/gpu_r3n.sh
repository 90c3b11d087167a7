#!/bin/bash
# GPU session r3n (1 GPU): GPU suite after the API additions (uncertainty mode fixture, sync_batch_norm option, transformers)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r3n_pytest.log 2>&1; echo "pytest exit $?"; tail -n 40 gpurun_out/r3n_pytest.log | cut -c1-400
