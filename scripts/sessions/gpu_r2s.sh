#!/bin/bash
# GPU session r2s (8 GPUs): weak-scaling bench line with the per-stage host timers of rank 0
mkdir -p gpurun_out
nproc; free -g | head -2
DCGC_PIPE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err; echo "bench n8 exit $?"
tail -2 gpurun_out/r2s_bench_n8.err
python - <<'PY'
import json
for l in open("gpurun_out/r2s_bench_n8.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("n=%d value %.0f ms %.3f e2e %.0f e2e_ms %.3f workers %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["host_workers"]))
        print(d["e2e"].get("pipe_trace_ms_per_step"))
PY
