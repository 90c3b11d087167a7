"""BASELINE configs 1 and 2 (parity-test cases, not the bench line), measured through the public API:
  cfg1  GraphConvModel regression on the real Delaney molecules (1 128 SMILES read by deepchem_b200/smiles.py,
        graph_conv_layers [64, 64], dense 128): seconds per epoch of fit() at the reference's batch size 100 and at
        one batch per epoch, final training RMSE, beside the CPU oracle port on the same molecules;
  cfg2  GraphConvModel 12-task classification on Tox21-shaped data (7 831 molecules, ~18.5 atoms, 2 classes, 25 % of
        the labels missing = zero weights): molecules/s of fit() (fwd + bwd + Adam, end to end from host memory) at
        batch 50 (the reference's Tox21 example), 1024 and 4096, beside the CPU oracle port.
Prints one JSON line per measurement."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch

dev = torch.device("cuda", 0)


def _fit_rate(model, ds, epochs, n):
    model.fit(ds, nb_epoch=1, deterministic=True)            # warm-up: allocator, layouts, kernels
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.fit(ds, nb_epoch=epochs, deterministic=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n * epochs / dt, dt / epochs


def _oracle_rate(mode, layers, n_tasks, pm, y, w, batch, steps=3):
    from oracle import graphconv_torch as O
    from oracle.convmol_layout import OracleConvMol, agglomerate, model_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    om = O.OracleGraphConvModel(n_tasks, layers, 128, mode=mode, batch_size=batch)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    om.train()
    mols = pm.slice(0, batch).to_list()
    yb = y[:batch]
    if mode == "classification":
        yb = np.eye(2, dtype=np.float32)[yb.astype(np.int64)]
    yt, wt = torch.from_numpy(yb), torch.from_numpy(w[:batch])
    times = []
    for it in range(steps + 1):
        t0 = time.perf_counter()
        mm = agglomerate([OracleConvMol(f, a) for f, a in mols])          # the reference rebuilds the layout every step
        inputs = [torch.from_numpy(np.asarray(a)) for a in model_inputs(mm)]
        inputs[0] = inputs[0].float()
        opt.zero_grad()
        loss = O.standard_loss(mode, om(inputs), yt, wt)
        loss.backward()
        opt.step()
        float(loss.detach())
        if it:
            times.append(time.perf_counter() - t0)
    return batch / min(times), torch.get_num_threads()


def cfg1():
    from deepchem_b200.data import CSVLoader
    from deepchem_b200.graphconvmodel import GraphConvModel
    ds = CSVLoader(["y"]).create_dataset(os.path.join(ROOT, "tests", "golden", "delaney.csv"))
    n = len(ds)
    for batch in (100, 1128):
        torch.manual_seed(0)
        m = GraphConvModel(1, [64, 64], 128, mode="regression", batch_size=batch, gemm_mode="tf32x3")
        rate, per_epoch = _fit_rate(m, ds, 30, n)
        rmse = float(np.sqrt(np.mean((m.predict(ds) - ds.y) ** 2)))
        print(json.dumps({"config": "cfg1 Delaney (real SMILES, 1128 molecules, 14991 atoms) GraphConv[64,64]+dense128 regression",
                          "batch": batch, "molecules_per_s": rate, "s_per_epoch": per_epoch, "epochs": 31,
                          "train_rmse_log_mol_per_l": rmse, "label_std": float(ds.y.std())}), flush=True)
    rate, threads = _oracle_rate("regression", [64, 64], 1, ds.X, ds.y, ds.w, 100)
    print(json.dumps({"config": "cfg1 CPU oracle port", "batch": 100, "molecules_per_s": rate, "cores": threads}), flush=True)


def cfg2():
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    n = 7831
    pm = make_molecules(n, seed=21, shape="tox21").pin_memory()
    y, w = make_labels(n, 12, "classification", seed=3, missing=0.25)
    ds = PackedDataset(pm, y, w)
    for batch in (50, 1024, 4096):
        torch.manual_seed(0)
        m = GraphConvModel(12, [64, 64], 128, mode="classification", n_classes=2, batch_size=batch, gemm_mode="tf32x3")
        rate, per_epoch = _fit_rate(m, ds, 5, n)
        print(json.dumps({"config": "cfg2 Tox21-shaped (7831 molecules, %d atoms) GraphConv[64,64]+dense128, 12 tasks x 2 classes, 25%% labels missing" % pm.n_atoms,
                          "batch": batch, "molecules_per_s": rate, "s_per_epoch": per_epoch,
                          "what": "GraphConvModel.fit end to end from host memory (layout build + H2D + fwd + bwd + Adam)"}), flush=True)
    for batch in (50, 1024):
        rate, threads = _oracle_rate("classification", [64, 64], 12, pm, y, w, batch)
        print(json.dumps({"config": "cfg2 CPU oracle port", "batch": batch, "molecules_per_s": rate, "cores": threads}), flush=True)


def cfg2_real():
    """Config 2 on the reference's own Tox21 file (tests/golden/tox21.csv.gz: 8 014 SMILES, 12 assays, 17 % of the
    labels missing) read by the RDKit-free reader."""
    from deepchem_b200.data import CSVLoader
    from deepchem_b200.graphconvmodel import GraphConvModel
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_tox21 import TASKS
    t0 = time.perf_counter()
    ds = CSVLoader(TASKS).create_dataset(os.path.join(ROOT, "tests", "golden", "tox21.csv.gz"))
    t_read = time.perf_counter() - t0
    ds.X.pin_memory()
    n = len(ds)
    for batch in (50, 1024):
        torch.manual_seed(0)
        m = GraphConvModel(12, [64, 64], 128, mode="classification", n_classes=2, batch_size=batch, gemm_mode="tf32x3")
        rate, per_epoch = _fit_rate(m, ds, 5, n)
        pred = m.predict(ds)
        y1 = np.eye(2, dtype=np.float32)[ds.y.astype(np.int64)]
        ce = float((-(y1 * np.log(np.clip(pred, 1e-12, 1.0))).sum(-1) * ds.w).sum() / ds.w.sum())
        print(json.dumps({"config": "cfg2 Tox21 (real SMILES, %d molecules, %d atoms) GraphConv[64,64]+dense128, 12 tasks x 2 classes" % (n, ds.X.n_atoms),
                          "batch": batch, "molecules_per_s": rate, "s_per_epoch": per_epoch, "epochs": 6,
                          "train_cross_entropy": ce, "smiles_read_s": t_read,
                          "what": "GraphConvModel.fit end to end from host memory (layout build + H2D + fwd + bwd + Adam)"}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["cfg1", "cfg2", "cfg2_real"]
    if "cfg2_real" in what:
        cfg2_real()
    if "cfg1" in what:
        cfg1()
    if "cfg2" in what:
        cfg2()
