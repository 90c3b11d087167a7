#!/bin/bash
# v4 (A operand in tensor memory) first light: per-GEMM fp64 tests, engine tests, A/B bench against v3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine_fp64.py tests/test_gpu_engine.py tests/test_gpu_dmpnn.py tests/test_gpu_mpnn.py -m gpu -q -x > gpurun_out/r4e_pytest.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/r4e_pytest.log | tail -n 3; grep -E "^FAILED|Error|^E  " gpurun_out/r4e_pytest.log | cut -c1-250 | head -n 30
for v in 0 1; do
DCGC_TC_V3=$v timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-e2e --sub "" --breakdown gpurun_out/r4e_breakdown_v3_$v.md > gpurun_out/r4e_bench_v3_$v.json 2> gpurun_out/r4e_bench_v3_$v.err; echo "bench v3=$v exit $?"; tail -n 3 gpurun_out/r4e_bench_v3_$v.err | cut -c1-300
python - <<P
import json
try:
    d = json.loads([l for l in open("gpurun_out/r4e_bench_v3_$v.json") if l.startswith("{")][-1])
    print("v3=$v value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]))
except Exception as e:
    print("no line", repr(e))
P
head -n 12 gpurun_out/r4e_breakdown_v3_$v.md
done
