#!/bin/bash
# GPU session r3q (1 GPU, last of the round): default bench at HEAD (after the lazy slab views), then the ncu launch list of the same command
mkdir -p gpurun_out
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/r3q_bench_n1.json 2> gpurun_out/r3q_bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r3q_bench_n1.json") if l.startswith("{")][-1])
    print("n=%d value %.0f ms %.4f e2e %.0f (%.4f ms) roof %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print("no bench line", e)
P
timeout 55 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3q_ncu_list.log 2>&1; echo "ncu list exit $?"
