#!/bin/bash
# GPU session r3h: everything that waits on a GPU slot, most important first:
#   full GPU suite at HEAD; the prefetch-race reproducer; shuffled-epoch rate; predict stage trace; bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r3h_pytest.log | cut -c1-200
for c in 0 1; do
  DBG_COMPACT=$c timeout 200 python scripts/debug_fit_crash.py > gpurun_out/r3h_race_compact$c.log 2>&1; echo "race reproducer (compact=$c) exit $?"; grep "fit call" gpurun_out/r3h_race_compact$c.log | tail -n 2
done
timeout 600 python scripts/shuffled_e2e.py > gpurun_out/r3h_shuffled.json 2> gpurun_out/r3h_shuffled.err; echo "shuffled exit $?"; cat gpurun_out/r3h_shuffled.json
timeout 300 python scripts/predict_trace.py > gpurun_out/r3h_predict.json 2> gpurun_out/r3h_predict.err; echo "predict exit $?"; cat gpurun_out/r3h_predict.json
timeout 600 python bench.py --breakdown gpurun_out/r3h_breakdown.md > gpurun_out/r3h_bench_n1.json 2> gpurun_out/r3h_bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r3h_bench_n1.json") if l.startswith("{")][-1])
print("value %.0f ms %.4f e2e %s roof %.3f cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]))
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3h_bench_reference_arm.json 2> gpurun_out/r3h_ref.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/r3h_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3h_ncu_list.log 2>&1; echo "ncu list exit $?"
