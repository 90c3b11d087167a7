#!/usr/bin/env python
"""Which ingredient of the small 'stress' regression case lifts the fused engine's gradient error above the fp32
oracle's?  Varies shape / mode / widths / batch-norm and the path (engine vs per-layer autograd ops)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import fp64_anchored_errors, oracle_batch, oracle_fp32_fp64
from oracle import graphconv_torch as O
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import make_labels, make_molecules


def run(shape, mode, layers, bn, path, gemm_mode="fp32", B=70, missing=0.25, n_tasks=3, seed=11, perturb=True):
    pm = make_molecules(B, seed=seed, shape=shape)
    y, w = make_labels(B, n_tasks, mode, seed=2, missing=missing)
    torch.manual_seed(0)
    om = O.OracleGraphConvModel(n_tasks, layers, 128, mode=mode, batch_size=B, batch_normalize=bn)
    if perturb:
        with torch.no_grad():
            for p in om.parameters():
                if p.dim() == 1:
                    p.add_(torch.randn_like(p) * 0.1)
    m = GraphConvModel(n_tasks, graph_conv_layers=layers, dense_layer_size=128, mode=mode, batch_size=B,
                       gemm_mode=gemm_mode, batch_normalize=bn, use_engine=(path == "engine"))
    m.model.load_state_dict(om.state_dict())
    batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
    inputs, labels, weights = m._prepare_batch(batch)
    if path == "engine":
        m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(), weights[0].contiguous(), B)
    else:
        m.model.train()
        outs = m.model(inputs)
        loss = m._loss_fn([outs[i] for i in m._loss_outputs], labels, weights)
        loss.backward()
    torch.cuda.synchronize()
    _, mm = oracle_batch(pm.to_list())
    res = oracle_fp32_fp64(om, mode, mm, B, batch[1][0], w)
    g32, g64 = res[torch.float32][2], res[torch.float64][2]
    rat, ours, base = [], [], []
    for n, p in m.model.named_parameters():
        if p.grad is None or float(g64[n].abs().max()) == 0:
            continue
        e, b, r, rb = fp64_anchored_errors(p.grad.detach().cpu(), g32[n], g64[n])
        ours.append(r); base.append(rb)
    print("%-7s %-14s %-16s bn=%d %-8s %-6s perturb=%d: rms ours median %.2e max %.2e | oracle median %.2e max %.2e" % (
        shape, mode, layers, bn, path, gemm_mode, perturb, np.median(ours), max(ours), np.median(base), max(base)), flush=True)


for shape in ("stress", "zinc"):
    for mode in ("regression", "classification"):
        run(shape, mode, [64, 64], True, "engine")
run("stress", "regression", [64, 64], True, "autograd")
run("stress", "regression", [64, 64], False, "engine")
run("stress", "regression", [64, 64], False, "autograd")
run("stress", "regression", [64, 64], True, "engine", perturb=False)
run("stress", "regression", [64, 64], True, "engine", missing=0.0)
run("stress", "regression", [64, 64], True, "engine", n_tasks=1)
run("stress", "regression", [128, 128, 128], True, "engine")
run("zinc", "regression", [128, 128, 128], True, "engine")
run("stress", "regression", [64, 64], True, "engine", B=512)
run("zinc", "regression", [64, 64], True, "engine", B=512)
