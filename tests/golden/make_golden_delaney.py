"""Delaney (ESOL) fixtures from the reference checkout.  Run in the build container:

    python tests/golden/make_golden_delaney.py

  delaney.csv            SMILES, the measured label, and the RDKit-derived descriptor columns the reference's own
                         ``datasets/delaney-processed.csv`` carries (molecular weight, H-bond donors, rings, minimum
                         degree): they pin hydrogen counts, ring counts and degrees of the RDKit-free SMILES reader
                         (deepchem_b200/smiles.py) over all 1 128 molecules.
  ref_delaney_real.npz   BASELINE config 1 on the real molecules: the first 100 Delaney molecules featurised by
                         deepchem_b200/smiles.py, then pushed through the REFERENCE's ConvMol / agglomerate_mols
                         (integer layout) and the reference's _GraphConvTorchModel (regression, [64, 64], dense 128,
                         batch 100: train-mode and eval-mode outputs, L2 loss against the measured labels).
"""
import csv
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/datasets/delaney-processed.csv"
DST = os.path.join(HERE, "delaney.csv")
COLS = [("smiles", "smiles"), ("measured log solubility in mols per litre", "y"), ("Molecular Weight", "mw"),
        ("Number of H-Bond Donors", "hbd"), ("Number of Rings", "rings"), ("Minimum Degree", "min_degree")]


def write_csv():
    with open(SRC) as fh:
        rows = list(csv.DictReader(fh))
    with open(DST, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([c for _, c in COLS])
        for r in rows:
            w.writerow([r[k].strip() for k, _ in COLS])
    print("%d molecules -> %s" % (len(rows), DST))
    return rows


def reference_run(rows, n=100):
    sys.path.insert(0, HERE)
    import make_golden as G            # imports the reference (rdkit stubbed) and the helpers
    import torch
    from deepchem.models.losses import L2Loss
    from deepchem.models.torch_models.graphconvmodel import _GraphConvTorchModel
    from deepchem_b200.smiles import atom_features, mol_from_smiles
    mols = []
    for r in rows[:n]:
        m = mol_from_smiles(r["smiles"].strip())
        mols.append((atom_features(m), m.adjacency_list()))
    cms, mm = G.ref_batch(mols)
    d = G.pack_mols(mols)
    d["features"] = d["features"].astype(np.float32)
    lay = G.layout_dict(cms, mm)
    for k in ("nodes", "mol_features_sorted"):
        lay[k] = lay[k].astype(np.float32)
    d.update({"ref_" + k: v for k, v in lay.items()})
    torch.manual_seed(6)
    model = _GraphConvTorchModel(1, graph_conv_layers=[64, 64], number_input_features=[75, 64], dense_layer_size=128,
                                 dropout=0.0, mode="regression", number_atom_features=75, n_classes=2,
                                 batch_normalize=True, uncertainty=False, batch_size=n)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    for k, v in model.state_dict().items():
        d["sd:" + k] = v.numpy().copy()
    args = G.layer_args(mm, n)
    model.train()
    res = model(args)
    for i, r in enumerate(res):
        d["ref_train_out%d" % i] = r.detach().numpy()
    y = np.asarray([float(r["measured log solubility in mols per litre"]) for r in rows[:n]], np.float32).reshape(n, 1)
    w = np.ones((n, 1), np.float32)
    losses = L2Loss()._create_pytorch_loss()(res[0], torch.from_numpy(y))
    d["y"], d["w"] = y, w
    d["ref_train_loss"] = (losses * torch.from_numpy(w)).mean().detach().numpy()
    model.eval()
    for i, r in enumerate(model(args)):
        d["ref_eval_out%d" % i] = r.detach().numpy()
    d["batch_size"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "ref_delaney_real.npz"), **d)
    print("reference run on %d real molecules: %d atoms, train loss %.6f" % (n, mm.get_num_atoms(),
                                                                             float(d["ref_train_loss"])))


if __name__ == "__main__":
    rows = write_csv()
    reference_run(rows)
