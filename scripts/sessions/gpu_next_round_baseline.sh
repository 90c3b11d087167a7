#!/bin/bash
# First GPU call of the next round (1 GPU, ~4 min): re-establish the baseline before changing anything.
#   gpurun --timeout 420 -- 'bash scripts/sessions/gpu_next_round_baseline.sh'
# Writes gpurun_out/r4a_*: GPU suite, smoke(), default bench (with cpu_baseline), reference arm, ncu launch list and ONE
# full ncu capture of the staged gather-sum (the roofline kernel of the bench line).
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r4a_pytest.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4a_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r4a_smoke.log
timeout 200 python bench.py > gpurun_out/r4a_bench_n1.json 2> gpurun_out/r4a_bench_n1.err; echo "bench exit $?"
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r4a_bench_reference_arm.json 2> gpurun_out/r4a_ref.err; echo "reference arm exit $?"
python - <<'P'
import json
for f in ("gpurun_out/r4a_bench_n1.json", "gpurun_out/r4a_bench_reference_arm.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"])
    except Exception as e:
        print(f, "no line", e)
P
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r4a_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mg_kernel -c 5 -o gpurun_out/r4a_mg_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r4a_ncu_full.log 2>&1; echo "ncu full exit $?"
