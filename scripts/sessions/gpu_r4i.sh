#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4i_pytest.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/r4i_pytest.log | tail -n 3; grep -E "^FAILED|Error|^E  " gpurun_out/r4i_pytest.log | cut -c1-250 | head -n 30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4i_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r4i_smoke.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --sub "" --breakdown gpurun_out/r4i_breakdown.md > gpurun_out/r4i_bench.json 2> gpurun_out/r4i_bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/r4i_bench.err | cut -c1-300
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r4i_bench.json") if l.startswith("{")][-1])
    print("value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"])
except Exception as e:
    print("no line", repr(e))
P
cat gpurun_out/r4i_breakdown.md
