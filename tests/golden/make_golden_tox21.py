"""Tox21 fixture from the reference checkout (BASELINE config 2 on the real molecules).  Run in the build container:

    python tests/golden/make_golden_tox21.py

  tox21.csv.gz   the SMILES and the 12 assay columns of the reference's ``datasets/tox21.csv.gz`` (8 014 rows, empty
                 cells = missing labels -> zero weights, ``data_loader.py:35-69``), ``mol_id`` dropped; written with a
                 fixed gzip mtime so the file is reproducible.
  ref_tox21_real.npz  BASELINE config 2 on the real molecules: rows 0..49 (the batch size of the reference's Tox21
                 example) featurised by deepchem_b200/smiles.py, pushed through the REFERENCE's ConvMol /
                 agglomerate_mols (integer layout) and the reference's _GraphConvTorchModel (classification, 12 tasks x
                 2 classes, [64, 64], dense 128): train-mode probabilities / logits / fingerprints, the reference's
                 SoftmaxCrossEntropy loss with zero weights on the missing labels reduced as _StandardLoss does
                 (torch_model.py:1267-1294), and the eval-mode outputs.
"""
import csv
import gzip
import io
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/datasets/tox21.csv.gz"
DST = os.path.join(HERE, "tox21.csv.gz")
TASKS = ['NR-AR', 'NR-AR-LBD', 'NR-AhR', 'NR-Aromatase', 'NR-ER', 'NR-ER-LBD', 'NR-PPAR-gamma', 'SR-ARE', 'SR-ATAD5',
         'SR-HSE', 'SR-MMP', 'SR-p53']

def reference_run(rows, n=50):
    import sys

    import numpy as np
    sys.path.insert(0, HERE)
    import make_golden as G            # imports the reference (rdkit stubbed) and the helpers
    import torch
    from deepchem.models.losses import SoftmaxCrossEntropy
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
    T = len(TASKS)
    torch.manual_seed(8)
    model = _GraphConvTorchModel(T, graph_conv_layers=[64, 64], number_input_features=[75, 64], dense_layer_size=128,
                                 dropout=0.0, mode="classification", number_atom_features=75, n_classes=2,
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
    y = np.zeros((n, T), np.float32)
    w = np.ones((n, T), np.float32)
    for i, r in enumerate(rows[:n]):
        for t, task in enumerate(TASKS):
            v = r[task].strip()
            if v == "":
                w[i, t] = 0.0
            else:
                y[i, t] = float(v)
    onehot = np.eye(2, dtype=np.float32)[y.astype(np.int64)]                  # graphconvmodel.py:404-406
    losses = SoftmaxCrossEntropy()._create_pytorch_loss()(res[1], torch.from_numpy(onehot))   # [n, T]
    d["y"], d["w"] = y, w
    d["ref_train_loss"] = (losses * torch.from_numpy(w)).mean().detach().numpy()
    model.eval()
    for i, r in enumerate(model(args)):
        d["ref_eval_out%d" % i] = r.detach().numpy()
    d["batch_size"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "ref_tox21_real.npz"), **d)
    print("reference run on %d real molecules: %d atoms, %d of %d labels missing, train loss %.6f"
          % (n, mm.get_num_atoms(), int((w == 0).sum()), w.size, float(d["ref_train_loss"])))


if __name__ == "__main__":
    with gzip.open(SRC, "rt", newline="") as fh:
        rows = list(csv.DictReader(fh))
    buf = io.StringIO()
    w = csv.writer(buf, lineterminator="\n")
    w.writerow(TASKS + ["smiles"])
    for r in rows:
        w.writerow([r[t].strip() for t in TASKS] + [r["smiles"].strip()])
    with open(DST, "wb") as out, gzip.GzipFile(fileobj=out, mode="wb", mtime=0, compresslevel=9) as gz:
        gz.write(buf.getvalue().encode())
    print("%d molecules -> %s (%d bytes)" % (len(rows), DST, os.path.getsize(DST)))
    reference_run(rows)
