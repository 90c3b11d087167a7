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
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "gather_sum_traffic.json")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per gather-sum launch from the committed ncu --set full capture
    (scripts/ncu_traffic.py writes profiles/gather_sum_traffic.json with the commit, the command and the sha256 of the
    kernel source it profiled).  Only quoted while that kernel source is unchanged; otherwise null."""
    import hashlib
    try:
        d = json.load(open(TRAFFIC_JSON))
        src = os.path.join(ROOT, "deepchem_b200", "csrc", "molgroup_kernels.cu")
        if hashlib.sha256(open(src, "rb").read()).hexdigest() != d["kernel_source_sha256"]:
            return None, "profiles/gather_sum_traffic.json is from an older molgroup_kernels.cu: not quoted"
        return float(d["traffic_bytes_per_launch"]), "profiles/gather_sum_traffic.json (commit %s, `%s`)" % (
            d.get("commit"), d.get("command"))
    except Exception as e:
        return None, "no capture on file (%s)" % type(e).__name__


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--gemm-mode", default=os.environ.get("DCGC_GEMM_MODE", "tf32x3"),
                    help="tf32x3 (default): tcgen05 tensor cores, 3-term TF32 split with fp32 accumulation, fp32-grade "
                         "results (same 1e-5 parity tests as fp32); fp32: SIMT FFMA")
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches rotated through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=200, help="the end-to-end leg runs max(steps, this) steps")
    ap.add_argument("--sub", default="dmpnn,predict", help="extra keyed sub-records after the headline (BASELINE configs "
                    "4 and 5): dmpnn, predict; '' for none")
    ap.add_argument("--predict-mols", type=int, default=1250000, help="molecules per GPU of the predict sub-record "
                    "(8 GPUs x 1.25 M = the 10 M of BASELINE configs[4])")
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
    """SM clock and clock-event (throttle) reasons of this rank's GPU, sampled every few milliseconds through NVML
    from the warm-up on; `window(t0, t1)` summarises the samples that fall inside a timed region
    (B200_PROFILING.md's clocks line).  Falls back to one nvidia-smi query per window when NVML is unavailable."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device, period=0.004):
        self.samples, self.period, self.stop_flag, self.thread, self.h, self.max_mhz = [], period, False, None, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0,1,2,3,4,5,6,7")
                                                               .split(",")[device.index or 0]))
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(reasons(self.h))))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1)

    def window(self, t0, t1):
        rows = [r for r in self.samples if t0 <= r[0] <= t1]
        if self.h is None or not rows:
            return self._smi()
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz, "samples": len(rows),
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "source": "nvml, %.0f ms period"
                % (self.period * 1e3)}

    def _smi(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0].split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": [],
                    "source": "nvidia-smi after the region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock source"]}


def make_pool(n_batches, batch, rank):
    from deepchem_b200.synthetic import make_labels, make_molecules
    pool = []
    for i in range(n_batches):
        # pinned + compact (exact int8 copy of the integer-valued features), as a packed shard on disk / in the e2e leg:
        # both legs then run the same arithmetic path (input_exact: the first layer skips the zero lo(A) term)
        pm = make_molecules(batch, seed=1000 * rank + i, shape="zinc").pin_memory()
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


def layout_build_cpu(batch, seed=0):
    """SURVEY 8(d) CPU line (3): ConvMol.agglomerate_mols (the oracle's restatement of the reference's Python, one
    core, per-molecule ConvMol objects prebuilt as at featurisation time) against the C++ layout builder (one core)."""
    from deepchem_b200.mol_graphs import BatchLayout
    from deepchem_b200.synthetic import make_molecules
    from oracle.convmol_layout import OracleConvMol, agglomerate
    pm = make_molecules(batch, seed=seed, shape="zinc")
    cms = [OracleConvMol(f, a) for f, a in pm.to_list()]
    t_py = []
    for _ in range(2):
        t0 = time.perf_counter()
        agglomerate(cms)
        t_py.append(time.perf_counter() - t0)
    t_c = []
    for _ in range(6):
        t0 = time.perf_counter()
        BatchLayout.build(pm, n_segments=batch)
        t_c.append(time.perf_counter() - t0)
    return {"agglomerate_mols_port": {"value": batch / min(t_py), "unit": "molecules/s", "cores": 1, "kind": "port",
                                      "sample": "one B=%d batch, best of 2" % batch},
            "cxx_layout_builder": {"value": batch / min(t_c[1:]), "unit": "molecules/s", "cores": 1,
                                   "sample": "one B=%d batch, best of 5 (dcgc_layout_plan + dcgc_layout_build: every slab "
                                             "section incl. CSR transpose and molecule groups)" % batch}}


def reference_as_is_line():
    """SURVEY 8(d) CPU line (1): the unmodified reference cannot run on the GPU box (no /root/reference there) and cannot
    run this configuration anywhere (widths fixed at 64, GraphConv detached, OOM at B=4096); its best smaller-batch
    number, measured in the build container by tests/golden/make_ref_cpu_throughput.py, is quoted from the fixture."""
    try:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cpu_throughput.json")))
        return {"value": d["best"]["molecules_per_s"], "unit": "molecules/s", "cores": d["cores"], "kind": "reference",
                "batch": d["best"]["batch"], "rows": d["rows"], "where": d["where"], "what": d["what"],
                "b4096": d["b4096"], "source": "tests/golden/ref_cpu_throughput.json (not measured in this run)"}
    except Exception as e:
        return {"unavailable": "fixture missing (%s)" % type(e).__name__}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = min(args.batch, 4096)
    steps, warmup = max(1, min(args.steps, 100)), max(0, min(args.warmup, 3))   # bounded: one step is ~1 s of 16 cores
    mol_s, sec, threads = oracle_cpu_throughput(batch, steps, warmup)
    line = {
        "impl": "reference", "metric": "GraphConv fwd+bwd molecules/sec", "value": mol_s, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": args.gpus * batch, "parallelism": "dp%d" % args.gpus,
                   "optimizer": "Adam (in step)",
                   "note": "CPU arm: every step is ONE B=%d batch on the host cores of rank 0 whatever --gpus says (a "
                           "bounded sample of the workload; molecules/s does not depend on how many are processed). "
                           "Oracle port (oracle/graphconv_torch.py + oracle/convmol_layout.py): the reference torch "
                           "model itself cannot run this config (width 64 hard-coded, GraphConv detached, OOM at "
                           "B=4096: SURVEY 0.3-0.5); its own best number is under cpu_baseline.lines" % batch},
        "cpu_baseline": {"value": mol_s, "unit": "molecules/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of one B=%d batch, layout rebuilt each step" % (steps, batch),
                         "lines": {"reference_as_is": reference_as_is_line()}},
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
    for pm, y, w in pool:
        batch = (model.batch_inputs(pm), [y], [w])
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

    sampler = ClockSampler(dev)          # NVML samples from the warm-up on; summarised per timed region below
    if rank == 0:
        sampler.start()
    for i in range(W):
        step_resident(i)
    barrier()

    # ---- timed region 1: device-resident (value); the dominant gather kernel is event-timed inside
    ops.profile_begin("dcgc_gather_sum")
    launches0 = ops.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record()
    for i in range(K):
        step_resident(W + i)
    e1.record()
    barrier()
    tw1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    clocks = sampler.window(tw0, tw1) if rank == 0 else None
    prof = ops.profile_end()
    value = world * B * K / (ms * 1e-3)
    # algorithmic bytes of the bracketed gather-sum launches, two ways:
    #  own    — what the fused kernel must move: forward 2*N*w*4 + E*4 (read X, write S, read the index list) with the
    #           76-float padded rows of the first layer; the backward launches also READ the self-path gradient they
    #           add to (the fused addend): 3*N*w*4 + E*4
    #  strict — SURVEY 8(d)'s unfused formulas: K1 = K5 = 2*n*w*4 + e*4, first-layer width 75
    widths_in = [76] + LAYERS[:-1]
    widths_strict = [75] + LAYERS[:-1]
    gs_bytes = gs_strict = 0
    for i in range(K):
        topo = resident[(W + i) % len(resident)][0][1]._dcgc_topology
        n_at, n_ed = topo.n_atoms, topo.n_edges
        gs_bytes += sum(2 * n_at * w_ * 4 + n_ed * 4 for w_ in widths_in)
        gs_bytes += sum(3 * n_at * w_ * 4 + n_ed * 4 for w_ in widths_in[1:])
        gs_strict += sum(2 * n_at * w_ * 4 + n_ed * 4 for w_ in widths_strict)
        gs_strict += sum(2 * n_at * w_ * 4 + n_ed * 4 for w_ in widths_strict[1:])

    # ---- per-scope pass (after the timed region, not part of it): the library's own CUDA-event scopes around every
    # kernel family for 5 more resident steps -> the `kernels` table of the JSON line (and --breakdown FILE).
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
        # ONE fit_generator call over warm-up + K2 batches (the pipeline is filled during the warm-up steps, as the
        # W warm-up steps of the resident region fill caches); the clock starts in the callback of the last
        # warm-up step, after a barrier + device synchronize, and stops after the same at the end.  K2 = max(K, 200):
        # the batches the pipeline already holds at t0 (reported as `prepared_at_t0`) are a few percent of the region.
        K2 = max(K, args.e2e_steps)
        warm_steps = 12
        mark = {}
        step0 = model._global_step

        def at_step(_model, step, **kw):
            if step - step0 == warm_steps:
                barrier()
                pf = getattr(model, "_active_prefetcher", None)
                mark["queued"] = pf.q.qsize() if pf is not None else 0
                mark["t0"] = time.perf_counter()

        gen = itertools.islice(model.default_generator(ds, epochs=100000, deterministic=True), warm_steps + K2)
        model.fit_generator(gen, checkpoint_interval=0, all_losses=losses, callbacks=[at_step])
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        ms2 = max_over_ranks((t_end - mark["t0"]) * 1e3)
        barrier()
        assert len(losses) == K2 + warm_steps
        e2e = {"value": world * B * K2 / (ms2 * 1e-3), "unit": "molecules/s", "ms_per_step": ms2 / K2, "steps": K2,
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "prepared_at_t0": {"queued_device_batches": mark.get("queued"), "layout_futures_at_most": model.host_workers,
                                  "in_preparation_at_most": 1,
                                  "note": "host work done before the clock started: at most this many of the %d timed "
                                          "batches" % K2},
               "clocks": sampler.window(mark["t0"], t_end) if rank == 0 else None,
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

    # ---- sub-records (BASELINE configs[3] and configs[4]): every rank takes part, rank 0 reports
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    subs = [s_ for s_ in args.sub.split(",") if s_]
    n_atoms_mean = sum(r[0][0].shape[0] for r in resident) / len(resident)
    host_workers = model.host_workers
    del resident, model
    torch.cuda.empty_cache()
    sub_records = {}
    ctx = dict(dev=dev, rank=rank, world=world, barrier=barrier, max_over_ranks=max_over_ranks, hbm_peak=hbm_peak,
               peak_src=peak_src, cpu=not args.no_cpu_baseline)
    if "dmpnn" in subs:
        sub_records["dmpnn"] = sub_dmpnn(args, ctx)
    if "predict" in subs:
        sub_records["predict"] = sub_predict(args, ctx)
    sampler.stop()
    if rank != 0:
        return
    roof = None
    if prof and prof["launches"]:
        achieved = gs_bytes / (prof["ms"] * 1e-3) / 1e9
        strict = gs_strict / (prof["ms"] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
        roof = {"bound": "hbm", "kernel": "mg_kernel<GatherSumOp> (molecule-group staged K1 neighbour gather-sum fwd + K5 transposed bwd)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "achieved_strict": strict, "frac_strict": strict / hbm_peak,
                "frac_of_nominal_8tbs": achieved / 8000.0, "frac_strict_of_nominal_8tbs": strict / 8000.0,
                "bytes_definition": {"frac": "the fused kernel's own compulsory bytes: fwd 2*N*w*4 + E*4 with 76-float first-"
                                             "layer rows, bwd 3*N*w*4 + E*4 (it also reads the self-path gradient it adds to)",
                                     "frac_strict": "SURVEY 8(d): K1 = K5 = 2*n*w*4 + e*4, first-layer width 75, no addend"},
                "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "launches": prof["launches"],
                "avg_launch_us": prof["ms"] * 1e3 / prof["launches"],
                "algorithmic_bytes_per_launch": gs_bytes / prof["launches"],
                "algorithmic_bytes_per_launch_strict": gs_strict / prof["launches"],
                "share_of_step": prof["ms"] / ms}
    cpu = None
    if not args.no_cpu_baseline:
        mol_s, sec, threads = oracle_cpu_throughput(B, 3, 1)
        cpu = {"value": mol_s, "unit": "molecules/s", "cores": threads, "kind": "port",
               "sample": "3 timed steps of one B=%d batch (oracle port, layout rebuilt each step)" % B,
               "lines": {"reference_as_is": reference_as_is_line(),
                         "oracle_port_exact_config": {"value": mol_s, "unit": "molecules/s", "cores": threads},
                         "layout_build": layout_build_cpu(B)}}
    line = {
        "metric": "GraphConv fwd+bwd molecules/sec", "value": value, "unit": "molecules/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "atoms_per_batch": n_atoms_mean,
                   "parallelism": "dp%d" % world, "gemm_mode": args.gemm_mode,
                   "arithmetic": "fp32 storage and accumulation everywhere; tf32x3 = every GEMM product as three tcgen05 "
                                 "MMAs over 11-bit operand halves (hi*hi + hi*lo + lo*hi), error ~2^-21, inside the 1e-5 "
                                 "parity bar (tests/test_gpu_engine_fp64.py pins this exact configuration to float64): "
                                 "kind::tf32 halves in the backward GEMMs, kind::f16 halves (same significand, K 16 per "
                                 "instruction) in the forward GEMMs, whose operands are activations inside fp16's range "
                                 "(sticky overflow flag checked after the run; DCGC_FWD_F16X3=0 = tf32 halves everywhere); "
                                 "fp32 = SIMT FFMA",
                   "optimizer": "Adam (in step)", "host_workers": host_workers,
                   "value_is": "kernel-only: device-resident batches, no layout build / H2D in the timed region; the "
                               "like-for-like number against the CPU arm is e2e",
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
    for k_, v_ in sub_records.items():
        line[k_] = v_
    # the forward GEMMs ran with fp16 operand halves: say whether any operand left fp16's range (0 = none did)
    from deepchem_b200 import _lib as _dl
    line["config"]["f16_overflow_flag"] = int(_dl.lib().dcgc_tc_f16_overflow())
    line["config"]["kernel_launches"] = ("programmatic dependent launch between the kernels of a step (DCGC_PDL=%s)"
                                         % os.environ.get("DCGC_PDL", "1"))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# sub-records: BASELINE configs[3] (D-MPNN) and configs[4] (sharded large-batch inference)
# ----------------------------------------------------------------------------------------------
def sub_dmpnn(args, ctx):
    """D-MPNN edge message passing, hidden 300, depth 3, QM9-shaped molecules, 12 targets, B = 4096 per GPU:
    fwd + L2 loss + bwd + Adam through the fused engine, data parallel (one all-reduce of the gradient slab)."""
    import torch
    from deepchem_b200 import ops
    from deepchem_b200.dmpnn import DMPNNModel, GraphDataset
    from deepchem_b200.dmpnn_data import make_graphs
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    B, K, W = args.batch, max(10, min(args.steps, 50)), 5
    pg = make_graphs(B, seed=1000 * rank, shape="qm9")
    y = np.random.default_rng(1000 * rank).standard_normal((B, 12)).astype(np.float32)
    torch.manual_seed(0)
    m = DMPNNModel(device=dev, n_tasks=12, batch_size=B, gemm_mode=args.gemm_mode)
    m.enable_data_parallel()
    ds = GraphDataset(pg, y)
    batch = next(m.default_generator(ds, deterministic=True))
    inputs, labels, weights = m._prepare_batch(batch)
    m.model.train()
    for _ in range(W):
        m._train_step(inputs, labels, weights)
    ctx["barrier"]()
    ops.profile_begin("dcgc_gather_sum")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        m._train_step(inputs, labels, weights)
    e1.record()
    ctx["barrier"]()
    ms = ctx["max_over_ranks"](e0.elapsed_time(e1))
    prof = ops.profile_end()
    # end to end: DMPNNModel.fit_generator over a pinned host shard (C++ table builder on worker threads, uploads +
    # f_ini assembly on a prefetch thread / side stream, step)
    big = make_graphs(4 * B, seed=1000 * rank + 1, shape="qm9").pin_memory()
    y4 = np.random.default_rng(1).standard_normal((4 * B, 12)).astype(np.float32)
    ds4 = GraphDataset(big, y4)
    m.fit_generator(m.default_generator(ds4, epochs=2, deterministic=True), checkpoint_interval=0)
    ctx["barrier"]()
    t0 = time.perf_counter()
    n_e2e = 40
    m.fit_generator(m.default_generator(ds4, epochs=n_e2e // 4, deterministic=True), checkpoint_interval=0)
    torch.cuda.synchronize()
    ms2 = ctx["max_over_ranks"]((time.perf_counter() - t0) * 1e3)
    ctx["barrier"]()
    if rank != 0:
        return None
    topo = inputs.topology
    info = topo.layout.info
    R, A, H, T = int(info.n_rows), int(info.n_atoms), 300, 3
    em, ea = int(info.n_map_entries), int(info.n_a2b_entries)
    # algorithmic bytes of the bond-row / atom gathers of one step: forward (T-1) message
    # gathers 2*R*H*4 + Em*4 and the atom aggregation (R + A)*H*4 + Ea*4; backward the transposed aggregation with the
    # ReLU mask folded in (A + 2R)*H*4 + Ea*4 and (T-1) transposed message gathers with mask and addend 4*R*H*4 + Em*4
    gb = (T - 1) * (2 * R * H * 4 + em * 4) + (R + A) * H * 4 + ea * 4 + (A + 2 * R) * H * 4 + ea * 4 + \
        (T - 1) * (4 * R * H * 4 + em * 4)
    rec = {"metric": "D-MPNN fwd+bwd molecules/sec", "value": world * B * K / (ms * 1e-3), "unit": "molecules/s",
           "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "scaling": "weak", "dtype": "f32",
           "engine": m._engine is not None,
           "config": {"workload": "qm9-synthetic B=%d/GPU (%d atoms, %d bond rows), atom 133 / bond 14 features, hidden "
                                  "300, depth 3, mean aggregation, FFN 300x3 -> 12 targets, regression, Adam in step"
                                  % (B, A, R), "global_batch": world * B, "parallelism": "dp%d" % world,
                      "gemm_mode": args.gemm_mode},
           "e2e": {"value": world * B * n_e2e / (ms2 * 1e-3), "unit": "molecules/s", "ms_per_step": ms2 / n_e2e,
                   "steps": n_e2e, "api": "DMPNNModel.fit_generator(default_generator(GraphDataset)) from pinned host memory"}}
    if prof and prof["launches"]:
        ach = gb * K / (prof["ms"] * 1e-3) / 1e9
        rec["roofline"] = {"bound": "hbm", "kernel": "fused_gather_kernel / gather_sum_kernel (CSR gathers over the two "
                                                     "D-MPNN index tables and their transposes)",
                           "achieved": ach, "peak": ctx["hbm_peak"], "unit": "GB/s", "frac": ach / ctx["hbm_peak"],
                           "peak_source": ctx["peak_src"], "traffic": None, "launches": prof["launches"],
                           "avg_launch_us": prof["ms"] * 1e3 / prof["launches"],
                           "algorithmic_bytes_per_step": gb, "share_of_step": prof["ms"] / ms}
    if ctx["cpu"]:
        rec["cpu_baseline"] = dmpnn_cpu_line(B)
    return rec


def dmpnn_cpu_line(B):
    import torch
    from deepchem_b200.dmpnn_data import make_graphs
    from oracle import dmpnn_torch as O
    torch.set_num_threads(os.cpu_count() or 1)
    pg = make_graphs(B, seed=0, shape="qm9")
    y = torch.from_numpy(np.random.default_rng(0).standard_normal((B, 12)).astype(np.float32))
    torch.manual_seed(0)
    om = O.OracleDMPNN(mode='regression', n_tasks=12)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    times = []
    for it in range(3):
        t0 = time.perf_counter()
        vals = [O.mapper_values(O.OracleGraph(*pg.graph(i)[:3])) for i in range(pg.n_mols)]   # the reference re-maps every batch
        batch = O.to_torch_batch(O.collate(vals))
        opt.zero_grad()
        loss = ((om(batch) - y) ** 2).mean()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    s_ = min(times[1:])
    return {"value": B / s_, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "best of 2 timed steps of one B=%d batch (oracle/dmpnn_torch.py: per-molecule mapper + collation + "
                      "encoder + FFN + Adam)" % B}


def sub_predict(args, ctx):
    """GraphConvModel.predict on PCBA-shaped molecules (128 tasks x 2 classes, GraphConv [64,64] + dense 128), each
    rank its contiguous shard of the dataset (predict(shard=(rank, world))), no collectives; host to host: layout
    build + H2D + forward + asynchronous D2H of the probabilities into one page-locked array."""
    import torch
    from deepchem_b200 import ops
    from deepchem_b200.data import PackedDataset, ReplayDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import PackedMols, make_molecules
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    B = args.batch
    per_gpu = max(B, args.predict_mols)
    shard = PackedMols.concat([make_molecules(B, seed=100 + i, shape="pcba") for i in range(8)]).pin_memory()
    total = per_gpu * world
    ds = ReplayDataset(shard, total)
    torch.manual_seed(0)
    m = GraphConvModel(128, [64, 64], 128, mode="classification", n_classes=2, batch_size=B, device=dev,
                       gemm_mode=args.gemm_mode)
    m.predict(PackedDataset(shard.slice(0, 2 * B)))                    # warm up (allocations, first launches)
    ctx["barrier"]()
    ops.profile_begin("dcgc_gather_sum")
    t0 = time.perf_counter()
    p = m.predict(ds, shard=(rank, world))
    torch.cuda.synchronize()
    sec = ctx["max_over_ranks"]((time.perf_counter() - t0) * 1e3) * 1e-3
    prof = ops.profile_end()
    ctx["barrier"]()
    assert p.shape == (per_gpu, 128, 2)
    if rank != 0:
        return None
    n_at, n_ed = shard.n_atoms / shard.n_mols * per_gpu, float(shard.adj_ptr[-1]) / shard.n_mols * per_gpu
    gb = sum(2 * n_at * w_ * 4 + n_ed * 4 for w_ in (76, 64))
    feat_b = 1 if getattr(shard, "features_i8", None) is not None else 4
    rec = {"metric": "GraphConvModel.predict molecules/sec (host to host)", "value": total / sec, "unit": "molecules/s",
           "n_gpus": world, "seconds": sec, "scaling": "weak", "dtype": "f32",
           "config": {"workload": "pcba-synthetic, %d molecules per GPU (%d in total; 8 GPUs = the 10 M of BASELINE "
                                  "configs[4]) streamed from a replayed %d-molecule pinned shard, 128 tasks x 2 classes, "
                                  "GraphConv[64,64]+dense128+BN, batch %d" % (per_gpu, total, shard.n_mols, B),
                      "parallelism": "%d contiguous shards, no collectives" % world, "gemm_mode": args.gemm_mode},
           "e2e": {"value": total / sec, "unit": "molecules/s",
                   "h2d_bytes_per_gpu": int(n_at * 75 * feat_b + n_ed * 4 * 3), "d2h_bytes_per_gpu": int(p.nbytes),
                   "api": "GraphConvModel.predict(ReplayDataset, shard=(rank, world))"}}
    if prof and prof["launches"]:
        ach = gb / (prof["ms"] * 1e-3) / 1e9
        rec["roofline"] = {"bound": "hbm", "kernel": "mg_kernel<GatherSumOp> (forward launches of the two conv layers)",
                           "achieved": ach, "peak": ctx["hbm_peak"], "unit": "GB/s", "frac": ach / ctx["hbm_peak"],
                           "peak_source": ctx["peak_src"], "traffic": None, "launches": prof["launches"],
                           "avg_launch_us": prof["ms"] * 1e3 / prof["launches"], "share_of_pass": prof["ms"] * 1e-3 / sec}
    if ctx["cpu"]:
        rec["cpu_baseline"] = predict_cpu_line(B)
    return rec


def predict_cpu_line(B):
    import torch
    from deepchem_b200.synthetic import make_molecules
    from oracle import graphconv_torch as O
    from oracle.convmol_layout import OracleConvMol, agglomerate, model_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    pm = make_molecules(B, seed=100, shape="pcba")
    mols = pm.to_list()
    torch.manual_seed(0)
    om = O.OracleGraphConvModel(128, [64, 64], 128, mode="classification", batch_size=B)
    om.eval()
    times = []
    with torch.no_grad():
        for _ in range(3):
            t0 = time.perf_counter()
            mm = agglomerate([OracleConvMol(f, a) for f, a in mols])
            inputs = [torch.from_numpy(np.asarray(a)) for a in model_inputs(mm)]
            inputs[0] = inputs[0].float()
            om(inputs)
            times.append(time.perf_counter() - t0)
    s_ = min(times[1:])
    return {"value": B / s_, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "best of 2 forward passes over one B=%d batch (oracle port, layout rebuilt each pass)" % B}


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
        # GraphGather forward: reads z, writes the fingerprint, the arg-max rows and (training) the two per-molecule
        # centred sums; backward: reads z and the per-molecule tensors, writes G = relu'(z) * BatchNorm-backward(dA)
        # directly (the dense layer's apply pass is inside this kernel since r6a)
        "dcgc_gather_fwd": act(D) + N * 4 + 5 * B * D * 4,
        "dcgc_gather_bwd": 2 * act(D) + N * 4 + 5 * B * D * 4,
        # conv forward GEMMs read [X | S] and write Y; the 4th call of the scope is the dense layer's input gradient
        "dcgc_group_gemm_fwd": sum(2 * act(w_) + act(c) for w_, c in zip(fp, LAYERS)) + act(D) + act(C),
        "dcgc_group_gemm_dgrad": sum(act(c) + 2 * act(w_) for w_, c in zip(fp[1:], LAYERS[1:])),
        "dcgc_group_gemm_wgrad": sum(2 * act(w_) + act(c) for w_, c in zip(fp, LAYERS)) + act(C) + act(D),
        "dcgc_linear_fwd": act(C) + act(D),
        "bn_relu_bwd_apply": sum(3 * act(c) for c in LAYERS),
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
