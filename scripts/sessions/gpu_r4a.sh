#!/bin/bash
# Round 2, call 1: GPU suite with the new fp64-anchored engine tests, the per-tensor parity probe, smoke(), default bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_engine_fp64.py > gpurun_out/r4a_pytest.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/r4a_pytest.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_engine_fp64.py -m gpu -q -s > gpurun_out/r4a_pytest_fp64.log 2>&1; echo "pytest fp64 exit $?"; grep -E "bench/|tox21/|passed|failed|Error" gpurun_out/r4a_pytest_fp64.log | cut -c1-250 | tail -n 30
timeout 600 python scripts/parity_probe.py > gpurun_out/r4a_parity_probe.json 2> gpurun_out/r4a_parity_probe.err; echo "probe exit $?"; cat gpurun_out/r4a_parity_probe.err | cut -c1-300 | tail -n 8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4a_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r4a_smoke.log | cut -c1-300
timeout 300 python bench.py --breakdown gpurun_out/r4a_breakdown.md > gpurun_out/r4a_bench_n1.json 2> gpurun_out/r4a_bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r4a_bench_n1.json") if l.startswith("{")][-1])
    print("value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
except Exception as e:
    print("no line", e)
P
