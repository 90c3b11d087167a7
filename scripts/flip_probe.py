#!/usr/bin/env python
"""Which gradient tensors of the 500-molecule unit-test batch (tests/test_gpu_tc.py::_engine_vs_float64) sit further
from float64 than 3x the fp32 oracle, in which GEMM mode, and how many atoms feed them: a per-degree weight of a bucket
with few atoms moves by O(1 / atoms of that degree) when ONE ReLU / max-pool decision flips.
python scripts/flip_probe.py [seed ...] > gpurun_out/flip_probe.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import fp64_anchored_errors, oracle_batch, oracle_fp32_fp64  # noqa: E402
from oracle import graphconv_torch as O  # noqa: E402
from deepchem_b200.data import PackedDataset  # noqa: E402
from deepchem_b200.graphconvmodel import GraphConvModel  # noqa: E402
from deepchem_b200.synthetic import make_labels, make_molecules  # noqa: E402


def main():
    seeds = [int(a) for a in sys.argv[1:]] or [5, 6, 7]
    rep = {}
    for seed in seeds:
        pm = make_molecules(500, seed=seed, shape="zinc")
        y, w = make_labels(500, 2, "regression", seed=1)
        torch.manual_seed(0)
        om = O.OracleGraphConvModel(2, [128, 128], 128, mode="regression", batch_size=500)
        _, mm = oracle_batch(pm.to_list())
        res = oracle_fp32_fp64(om, "regression", mm, 500, y, w)
        _, _, g64 = res[torch.float64]
        _, _, g32 = res[torch.float32]
        deg_count = np.diff(np.asarray(mm.deg_slice)[:, 0].tolist() + [mm.get_atom_features().shape[0]]) \
            if hasattr(mm, "deg_slice") else None
        for mode in ("tf32x3", "fp32"):
            m = GraphConvModel(2, [128, 128], 128, mode="regression", batch_size=500, gemm_mode=mode)
            m.model.load_state_dict(om.state_dict())
            batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
            inputs, labels, weights = m._prepare_batch(batch)
            m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0], weights[0], 500)
            rows = {}
            for name, p in m.model.named_parameters():
                e, base, r, rbase = fp64_anchored_errors(p.grad.detach().cpu(), g32[name], g64[name])
                if e > max(1e-5, 3 * base) or r > max(1e-5, 3 * rbase):
                    rows[name] = [e, base, r, rbase]
            rep["seed%d/%s" % (seed, mode)] = rows
            sys.stderr.write("seed %d %s: %d tensors outside 3x: %s\n" % (seed, mode, len(rows), json.dumps(rows)))
        rep["seed%d/deg_slice" % seed] = np.asarray(mm.deg_slice).tolist()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
