#!/usr/bin/env python
"""Benchmark of the GraphConv hot path (BASELINE.json metric: GraphConv fwd+bwd molecules/sec at
1/2/4/8 B200; achieved HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port on host cores

Workload (configs[2] of BASELINE.json, the configuration the metric is quoted on): synthetic
ZINC-shaped ConvMol stream (25 avg heavy atoms, 75-dim features), batch 4096 per GPU, GraphConv
[128,128,128], dense 128, BatchNorm on, regression, Adam step included.  A "step" is one
fwd+bwd+optimizer pass over one batch per GPU (weak scaling: per-GPU batch fixed).

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same metric
through the public API (GraphConvModel.fit_on_batch on a host PackedMols shard: C++ layout build,
H2D from pinned memory, step, D2H of the loss) every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "zinc-synthetic B=4096/GPU, 25 atoms avg, F=75, GraphConv[128,128,128]+dense128+BN, regression T=1"
LAYERS = [128, 128, 128]
DENSE = 128
# dram__bytes_read.sum + dram__bytes_write.sum per gather-sum launch, mean over the 5 launches of one step
# (layer-0 forward 34.9 MB, 2 forward 61.1 / 63.0 MB, 2 backward 116.8 / 118.4 MB), from the ncu --set full
# capture of this command summarised in profiles/r2f_ncu_staged_kernels.md: BELOW the algorithmic bytes because
# most of the 52 MB a launch writes is still in the 126 MB L2 when the kernel ends (DRAM writes 0.5-9 MB).
NCU_GATHER_SUM_TRAFFIC = 78.8e6
NCU_TRAFFIC_SOURCE = ("profiles/r2f_ncu_staged_kernels.md (ncu --set full, dram__bytes_read.sum + "
                      "dram__bytes_write.sum, mean of the 5 launches of a step)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--gemm-mode", default=os.environ.get("DCGC_GEMM_MODE", "tf32x3"),
                    help="tf32x3 (default): tcgen05 tensor cores, 3-term TF32 split with fp32 accumulation, fp32-grade "
                         "results (same 1e-5 parity tests as fp32); fp32: SIMT FFMA")
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches rotated through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write an in-situ per-scope CUDA-event breakdown "
                    "(5 extra steps, outside the timed region) to this file")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# algorithmic bytes: fp32, each operand touched once
# ----------------------------------------------------------------------------------------------
def gather_sum_bytes(n_rows, n_edges, width):
    """read X once + write S once + read the index list"""
    return 2 * n_rows * width * 4 + n_edges * 4


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_pool(n_batches, batch, rank):
    from deepchem_b200.synthetic import make_labels, make_molecules
    pool = []
    for i in range(n_batches):
        pm = make_molecules(batch, seed=1000 * rank + i, shape="zinc")
        y, w = make_labels(batch, 1, "regression", seed=1000 * rank + i)
        pool.append((pm, y, w))
    return pool


# ----------------------------------------------------------------------------------------------
# CPU arm: oracle port timed on the host cores
# ----------------------------------------------------------------------------------------------
def oracle_cpu_throughput(batch, steps, warmup, seed=0):
    import torch
    from deepchem_b200.synthetic import make_labels, make_molecules
    from oracle import graphconv_torch as O
    from oracle.convmol_layout import OracleConvMol, agglomerate, model_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    pm = make_molecules(batch, seed=seed, shape="zinc")
    y, w = make_labels(batch, 1, "regression", seed=seed)
    mols = pm.to_list()
    torch.manual_seed(0)
    om = O.OracleGraphConvModel(1, LAYERS, DENSE, mode="regression", batch_size=batch)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    om.train()
    yt, wt = torch.from_numpy(y), torch.from_numpy(w)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        # the reference rebuilds the batch layout every step (graphconvmodel.py:414)
        mm = agglomerate([OracleConvMol(f, a) for f, a in mols])
        inputs = [torch.from_numpy(np.asarray(a)) for a in model_inputs(mm)]
        inputs[0] = inputs[0].float()
        opt.zero_grad()
        out = om(inputs)
        loss = O.standard_loss("regression", out, yt, wt)
        loss.backward()
        opt.step()
        float(loss.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = min(args.batch, 4096)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    mol_s, sec, threads = oracle_cpu_throughput(batch, steps, warmup)
    line = {
        "impl": "reference", "metric": "GraphConv fwd+bwd molecules/sec", "value": mol_s, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": batch,
                   "note": "CPU oracle port (oracle/graphconv_torch.py + oracle/convmol_layout.py); the "
                           "reference torch model itself cannot run this config (width 64 hard-coded, "
                           "GraphConv detached, OOM at B=4096: SURVEY 0.3-0.5)"},
        "cpu_baseline": {"value": mol_s, "unit": "molecules/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of one B=%d batch, layout rebuilt each step" % (steps, batch)},
        "e2e": {"value": mol_s, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepchem_b200 import _lib, ops, parallel
    from deepchem_b200.graphconvmodel import GraphConvModel

    rank, world, local = parallel.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if _lib.lib().dcgc_device_ok() != 1:
        raise SystemExit("libdcgc: no sm_100 device visible")
    B, K, W = args.batch, args.steps, max(3, args.warmup)

    torch.manual_seed(0)
    model = GraphConvModel(1, LAYERS, DENSE, mode="regression", batch_size=B, device=dev,
                           gemm_mode=args.gemm_mode)
    model.enable_data_parallel()
    model.model.train()
    pool = make_pool(args.pool, B, rank)

    # ---- device-resident inputs for `value`
    resident = []
    h2d_bytes = 0
    for pm, y, w in pool:
        batch = (model.batch_inputs(pm), [y], [w])
        h2d_bytes = int(batch[0].layout.info.slab_bytes) + pm.features.nbytes + y.nbytes + w.nbytes
        resident.append(model._prepare_batch(batch))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def step_resident(i):
        return model._train_step(*resident[i % len(resident)])

    for i in range(W):
        step_resident(i)
    barrier()

    # ---- timed region 1: device-resident (value); the dominant gather kernel is event-timed inside
    ops.profile_begin("dcgc_gather_sum")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step_resident(W + i)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    prof = ops.profile_end()
    value = world * B * K / (ms * 1e-3)
    # algorithmic bytes of the bracketed gather-sum launches: per layer forward
    # 2*N*w*4 + E*4 (read X once, write S once, read the index list); the backward launches also
    # read the self-path gradient they add to: 3*N*w*4 + E*4
    widths_in = [76] + LAYERS[:-1]
    gs_bytes = 0
    for i in range(K):
        topo = resident[(W + i) % len(resident)][0][1]._dcgc_topology
        n_at, n_ed = topo.n_atoms, topo.n_edges
        gs_bytes += sum(2 * n_at * w_ * 4 + n_ed * 4 for w_ in widths_in)
        gs_bytes += sum(3 * n_at * w_ * 4 + n_ed * 4 for w_ in widths_in[1:])

    # ---- per-scope pass (rank 0, after the timed region, not part of it): the library's own CUDA-event scopes around
    # every kernel family for 5 more resident steps -> the `kernels` table of the JSON line (and --breakdown FILE)
    # EVERY rank runs the extra steps (a data-parallel step holds a collective: ranks must stay in lockstep); only
    # rank 0 brackets them with the event scopes
    kernels = None
    nb = 5
    if rank == 0:
        ops.profile_begin("*")
    for i in range(nb):
        step_resident(i)
    torch.cuda.synchronize()
    if rank == 0:
        rep = ops.profile_report()
        tot = sum(v[0] for v in rep.values())
        kernels = kernel_table(rep, nb, [r[0][1]._dcgc_topology for r in resident], B)
    if args.breakdown and rank == 0:
        with open(args.breakdown, "w") as fh:
            fh.write("# in-situ CUDA-event breakdown, %d steps, %s, ms_per_step(timed)=%.4f\n" % (nb, args.gemm_mode, ms / K))
            fh.write("| scope | us/step | share | calls/step | us/call |\n|---|---:|---:|---:|---:|\n")
            for name, (t, n) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
                fh.write("| %s | %.1f | %.1f%% | %.1f | %.1f |\n" % (name, t * 1e3 / nb, 100 * t / tot, n / nb, t * 1e3 / n))
            fh.write("| total | %.1f | | | |\n" % (tot * 1e3 / nb))
    barrier()

    # ---- timed region 2: end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        import itertools
        from deepchem_b200.data import PackedDataset
        from deepchem_b200.synthetic import PackedMols
        big = PackedMols.concat([pm for pm, _, _ in pool]).pin_memory()
        from deepchem_b200 import graphconvmodel as G
        feat_item = 1 if (big.features_i8 is not None and G._USE_I8) else 4   # exact int8 copy of 0/1 features
        topos = [r[0][1]._dcgc_topology for r in resident]
        h2d_bytes = int(np.mean([int(t_.layout.info.slab_bytes) + t_.n_atoms * pool[0][0].n_feat * feat_item
                                 for t_ in topos])) + pool[0][1].nbytes + pool[0][2].nbytes
        ds = PackedDataset(big, np.concatenate([y for _, y, _ in pool]), np.concatenate([w for _, _, w in pool]))
        model.log_frequency = 1                                # loss read back to the host every step
        losses = []
        # ONE fit_generator call over warm-up + K batches (the pipeline is filled during the warm-up steps, as the
        # W warm-up steps of the resident region fill caches); the clock starts in the callback of the last
        # warm-up step, after a barrier + device synchronize, and stops after the same at the end
        warm_steps = 12
        mark = {}
        step0 = model._global_step

        def at_step(_model, step, **kw):
            if step - step0 == warm_steps:
                barrier()
                mark["t0"] = time.perf_counter()

        gen = itertools.islice(model.default_generator(ds, epochs=1000, deterministic=True), warm_steps + K)
        model.fit_generator(gen, checkpoint_interval=0, all_losses=losses, callbacks=[at_step])
        torch.cuda.synchronize()
        ms2 = max_over_ranks((time.perf_counter() - mark["t0"]) * 1e3)
        barrier()
        assert len(losses) == K + warm_steps
        e2e = {"value": world * B * K / (ms2 * 1e-3), "unit": "molecules/s", "ms_per_step": ms2 / K,
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "feature_upload": "int8 (exact copy of the integer-valued feature matrix kept by the packed shard)"
                                 if feat_item == 1 else "fp32",
               "api": "GraphConvModel.fit_generator(default_generator(PackedDataset)) with log_frequency=1: "
                      "C++ layout build (worker threads) + H2D from pinned host memory + fwd/bwd/Adam + a 4-byte "
                      "loss readback for every step (asynchronous, consumed one step late); host work of the "
                      "next batches overlaps the GPU on a prefetch thread"}

    if e2e is not None and getattr(model, "_pipe_trace", None) is not None:
        # DCGC_PIPE_TRACE=1: host-side stage timers of the fit pipeline of this rank (ms per step, incl. warm-up steps)
        tr = model._pipe_trace
        n_tr = max(1.0, tr.get("fit_steps", 1.0))
        e2e["pipe_trace_ms_per_step"] = {k: round(v / n_tr * 1e3, 4) for k, v in tr.items()
                                         if k not in ("fit_steps", "pf_batches")}
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roof = None
    if prof and prof["launches"]:
        prof["bytes"] = gs_bytes
        achieved = prof["bytes"] / (prof["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "mg_kernel<GatherSumOp> (molecule-group staged K1 neighbour gather-sum fwd + K5 transposed bwd)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src, "traffic": NCU_GATHER_SUM_TRAFFIC, "traffic_source": NCU_TRAFFIC_SOURCE,
                "launches": prof["launches"],
                "avg_launch_us": prof["ms"] * 1e3 / prof["launches"],
                "algorithmic_bytes_per_launch": prof["bytes"] / prof["launches"],
                "share_of_step": prof["ms"] / ms}
    cpu = None
    if not args.no_cpu_baseline:
        mol_s, sec, threads = oracle_cpu_throughput(B, 3, 1)
        cpu = {"value": mol_s, "unit": "molecules/s", "cores": threads, "kind": "port",
               "sample": "3 timed steps of one B=%d batch (oracle port, layout rebuilt each step)" % B}
    n_atoms = sum(r[0][0].shape[0] for r in resident) / len(resident)
    line = {
        "metric": "GraphConv fwd+bwd molecules/sec", "value": value, "unit": "molecules/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "atoms_per_batch": n_atoms,
                   "parallelism": "dp%d" % world, "gemm_mode": args.gemm_mode,
                   "arithmetic": "fp32 storage and accumulation everywhere; tf32x3 = every GEMM product as three tcgen05 "
                                 "kind::tf32 MMAs (hi*hi + hi*lo + lo*hi), error ~2^-21, inside the 1e-5 parity bar "
                                 "(tests/test_gpu_tc.py); fp32 = SIMT FFMA",
                   "optimizer": "Adam (in step)", "host_workers": model.host_workers,
                   "l2": "no explicit flush: %d distinct batches rotated, >1 GB touched per step (> 126 MB L2)"
                         % len(pool)},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    if kernels is not None:
        for row in kernels:
            if row["algorithmic_mb_per_step"] is not None:
                row["achieved_gbs"] = row["algorithmic_mb_per_step"] * 1e-3 / (row["us_per_step"] * 1e-6)
                row["frac_of_hbm_peak"] = row["achieved_gbs"] / hbm_peak
        line["kernels"] = {"source": "5 extra resident steps after the timed region with the library's CUDA-event scopes "
                                     "around every kernel family (rank 0); bytes = activation operands touched once "
                                     "(DESIGN.md section 3), weights and index tables of the GEMMs not counted",
                           "rows": kernels}
    print(json.dumps(line), flush=True)


def kernel_table(rep, n_steps, topos, B):
    """Per kernel family of one training step: in-situ time and the algorithmic bytes DESIGN.md section 3 states for
    it (mean over the rotated batches), so every family's distance from the HBM roofline is on the bench line, not
    only the gather-sum's."""
    N = float(np.mean([t.n_atoms for t in topos]))
    E = float(np.mean([t.n_edges for t in topos]))
    fp = [76] + LAYERS[:-1]                       # padded input width of each conv layer
    C, D, L = LAYERS[-1], DENSE, len(LAYERS)
    act = lambda w_: N * w_ * 4                   # noqa: E731  one fp32 activation matrix of width w_
    bytes_per_step = {
        "dcgc_gather_sum": sum(2 * act(w_) + E * 4 for w_ in fp) + sum(3 * act(w_) + E * 4 for w_ in fp[1:]),
        "dcgc_pool_fwd": sum(2 * act(c) + E * 4 + N * c for c in LAYERS),
        "dcgc_pool_bwd": sum(3 * act(c) + 2 * E * 4 + N * c for c in LAYERS),
        "dcgc_gather_fwd": act(D) + N * 4 + 2 * B * D * 4,
        "dcgc_gather_bwd": act(D) + N * 4 + 2 * B * D * 4,
        # conv forward GEMMs read [X | S] and write Y; the 4th call of the scope is the dense layer's input gradient
        "dcgc_group_gemm_fwd": sum(2 * act(w_) + act(c) for w_, c in zip(fp, LAYERS)) + act(D) + act(C),
        "dcgc_group_gemm_dgrad": sum(act(c) + 2 * act(w_) for w_, c in zip(fp[1:], LAYERS[1:])),
        "dcgc_group_gemm_wgrad": sum(2 * act(w_) + act(c) for w_, c in zip(fp, LAYERS)) + act(C) + act(D),
        "dcgc_linear_fwd": act(C) + act(D),
        "bn_relu_bwd_apply": sum(3 * act(c) for c in LAYERS) + 3 * act(D),
    }
    rows = []
    for name, (t_ms, n_calls) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
        b = bytes_per_step.get(name)
        rows.append({"scope": name, "us_per_step": t_ms * 1e3 / n_steps, "calls_per_step": n_calls / n_steps,
                     "algorithmic_mb_per_step": None if b is None else b / 1e6})
    return rows


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
