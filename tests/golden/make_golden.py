"""Generate golden fixtures by running the REFERENCE (pandegroup/deepchem at /root/reference)
in the build container.  The fixtures (npz) are committed; this script documents how.

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  kat_ccc_c.npz        the reference's own known-answer vectors for the path
                       (deepchem/models/tests/assets/*.npy used by
                       models/tests/test_layers.py:1458-1546 and
                       models/tests/test_graphconv_torchmodel.py:15-95) with the rdkit-free
                       reconstruction of the ['CCC','C'] ConvMol inputs, replayed through the
                       reference layers to confirm the reconstruction.
  ref_layout_*.npz     reference ConvMol / agglomerate_mols integer outputs on seeded batches
  ref_layers.npz       reference GraphConv / GraphPool / GraphGather forward on a seeded batch
  ref_model_*.npz      reference _GraphConvTorchModel forward (+ loss) on a seeded batch
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _refimport import import_reference  # noqa: E402

dc = import_reference()
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import deepchem.models.torch_models.layers as L  # noqa: E402
from deepchem.feat.mol_graphs import ConvMol  # noqa: E402
from deepchem.models.torch_models.graphconvmodel import _GraphConvTorchModel  # noqa: E402
from deepchem.models.losses import L2Loss, SoftmaxCrossEntropy  # noqa: E402
from deepchem_b200.synthetic import make_molecules  # noqa: E402

ASSETS = "/root/reference/deepchem/models/tests/assets/"


def carbon(deg, implicit_valence, n_h):
    """75-dim ConvMolFeaturizer row of an sp3 carbon (feat/graph_features.py:322-391)."""
    f = np.zeros(75)
    f[0] = 1
    f[44 + deg] = 1
    f[55 + implicit_valence] = 1
    f[66] = 1
    f[70 + n_h] = 1
    return f


def pack_mols(mols):
    """list of (features, adj) -> flat arrays storable in npz."""
    atom_ptr = np.cumsum([0] + [len(a) for _, a in mols]).astype(np.int32)
    adj_ptr, adj_idx = [0], []
    for _, adj in mols:
        for nb in adj:
            adj_idx.extend(nb)
            adj_ptr.append(len(adj_idx))
    feats = np.concatenate([np.asarray(f, dtype=np.float64).reshape(len(a), -1) for f, a in mols])
    return dict(atom_ptr=atom_ptr, adj_ptr=np.asarray(adj_ptr, np.int32),
                adj_idx=np.asarray(adj_idx, np.int32), features=feats)


def ref_batch(mols):
    cms = [ConvMol(np.asarray(f, dtype=np.float64), adj) for f, adj in mols]
    return cms, ConvMol.agglomerate_mols(cms)


def layer_args(mm, n_samples=None):
    x = torch.from_numpy(mm.get_atom_features().astype(np.float32))
    args = [x, torch.from_numpy(mm.deg_slice), torch.from_numpy(mm.membership)]
    if n_samples is not None:
        args.append(torch.tensor(n_samples))
    return args + [torch.from_numpy(a) for a in mm.get_deg_adjacency_lists()[1:]]


def layout_dict(cms, mm):
    d = dict(deg_slice=mm.deg_slice, membership=mm.membership, nodes=mm.get_atom_features())
    for k, a in enumerate(mm.get_deg_adjacency_lists()):
        d["deg_adj_%d" % k] = a
    # per-molecule ConvMol outputs (a1)
    d["mol_deg_slice"] = np.stack([c.deg_slice for c in cms])
    d["mol_deg_block_indices"] = np.concatenate([c.deg_block_indices for c in cms])
    d["mol_degree_list"] = np.concatenate([np.asarray(c.degree_list, np.int32) for c in cms])
    d["mol_canon_adj_flat"] = np.asarray(
        [k for c in cms for nb in c.get_adjacency_list() for k in nb], np.int32)
    d["mol_features_sorted"] = np.concatenate([c.get_atom_features() for c in cms])
    return d


def main():
    torch.manual_seed(0)
    # ---------------------------------------------------------------- KATs (CCC + C)
    mols = [(np.stack([carbon(1, 3, 3), carbon(2, 2, 2), carbon(1, 3, 3)]), [[1], [0, 2], [1]]),
            (np.stack([carbon(0, 4, 4)]), [[]])]
    cms, mm = ref_batch(mols)
    out = pack_mols(mols)
    names = ["graphconvlayer_weights", "graphconvlayer_biases", "graphconvlayer_result",
             "graphpoollayer_result", "graphgatherlayer_result",
             "graphconvlayer0_weights", "graphconvlayer0_biases",
             "graphconvlayer1_weights", "graphconvlayer1_biases",
             "dense_weights", "dense_biases", "reshapedense_weights", "reshapedense_biases",
             "graphconvmodel_output_classification", "graphconvmodel_logits_classification",
             "graphconvmodel_neural_classification"]
    A = {n: np.load(ASSETS + n + ".npy", allow_pickle=True) for n in names}
    for n in names:
        out["asset_" + n] = np.asarray(A[n].tolist(), dtype=np.float32)
    # replay through the reference to confirm the reconstruction of the inputs
    args = layer_args(mm)
    conv = L.GraphConv(2, number_input_features=75)
    conv.W_list = nn.ParameterList([nn.Parameter(torch.tensor(k)) for k in A["graphconvlayer_weights"].tolist()])
    conv.b_list = nn.ParameterList([nn.Parameter(torch.tensor(k)) for k in A["graphconvlayer_biases"].tolist()])
    errs = [np.abs(conv(args).detach().numpy() - A["graphconvlayer_result"]).max(),
            np.abs(L.GraphPool()(args).numpy() - A["graphpoollayer_result"]).max(),
            np.abs(L.GraphGather(2)(args).numpy() - A["graphgatherlayer_result"]).max()]
    model = _GraphConvTorchModel(2, graph_conv_layers=[64, 64], number_input_features=[75, 64],
                                 dense_layer_size=128, dropout=0.0, mode="classification",
                                 number_atom_features=75, n_classes=2, batch_normalize=False,
                                 uncertainty=False, batch_size=10)
    for i in (0, 1):
        model.graph_convs[i].W_list = nn.ParameterList(
            [nn.Parameter(torch.tensor(k)) for k in A["graphconvlayer%d_weights" % i].tolist()])
        model.graph_convs[i].b_list = nn.ParameterList(
            [nn.Parameter(torch.tensor(k)) for k in A["graphconvlayer%d_biases" % i].tolist()])
    model.dense.weight.data = torch.from_numpy(np.transpose(A["dense_weights"]))
    model.dense.bias.data = torch.from_numpy(A["dense_biases"])
    model.reshape_dense.weight.data = torch.from_numpy(np.transpose(A["reshapedense_weights"]))
    model.reshape_dense.bias.data = torch.from_numpy(A["reshapedense_biases"])
    res = model(layer_args(mm, 2))
    errs += [np.abs(res[0].detach().numpy() - A["graphconvmodel_output_classification"]).max(),
             np.abs(res[1].detach().numpy() - A["graphconvmodel_logits_classification"]).max(),
             np.abs(res[2].detach().numpy() - A["graphconvmodel_neural_classification"]).max()]
    print("KAT replay max-abs errors (conv, pool, gather, output, logits, neural):", errs)
    assert max(errs) < 1e-5
    out.update({"ref_" + k: v for k, v in layout_dict(cms, mm).items()})
    np.savez_compressed(os.path.join(HERE, "kat_ccc_c.npz"), **out)

    # ---------------------------------------------------------------- layout goldens
    for name, shape, n, seed in (("stress", "stress", 48, 11), ("zinc", "zinc", 64, 12),
                                 ("delaney", "delaney", 33, 13)):
        pm = make_molecules(n, seed=seed, shape=shape)
        mols = pm.to_list()
        cms, mm = ref_batch(mols)
        d = pack_mols(mols)
        d["features"] = d["features"].astype(np.float32)
        lay = layout_dict(cms, mm)
        for k in ("nodes", "mol_features_sorted"):
            lay[k] = lay[k].astype(np.float32)
        d.update({"ref_" + k: v for k, v in lay.items()})
        np.savez_compressed(os.path.join(HERE, "ref_layout_%s.npz" % name), **d)
        print("layout", name, "mols", n, "atoms", mm.get_num_atoms(),
              "deg sizes", mm.deg_slice[:, 1].tolist())

    # the null molecule (one atom of each degree bonded to itself, mol_graphs.py:236-254)
    np.random.seed(5)
    nm = ConvMol.get_null_mol(7)
    mm = ConvMol.agglomerate_mols([nm, nm])
    d = dict(features=nm.get_atom_features(), deg_slice=mm.deg_slice, membership=mm.membership)
    for k, a in enumerate(mm.get_deg_adjacency_lists()):
        d["deg_adj_%d" % k] = a
    np.savez_compressed(os.path.join(HERE, "ref_layout_nullmol.npz"), **d)

    # ---------------------------------------------------------------- layer forward goldens
    pm = make_molecules(40, seed=21, shape="stress", dense_features=True)
    mols = pm.to_list()
    cms, mm = ref_batch(mols)
    d = pack_mols(mols)
    d["features"] = d["features"].astype(np.float32)
    args = layer_args(mm)
    torch.manual_seed(3)
    conv = L.GraphConv(64, number_input_features=75, activation_fn=torch.relu)
    with torch.no_grad():
        for b in conv.b_list:
            b.normal_(0, 0.3)
    d["W"] = np.stack([w.detach().numpy() for w in conv.W_list])
    d["b"] = np.stack([b.detach().numpy() for b in conv.b_list])
    d["ref_conv_relu"] = conv(args).detach().numpy()
    conv.activation_fn = None
    d["ref_conv_linear"] = conv(args).detach().numpy()
    d["ref_pool"] = L.GraphPool()(args).numpy()
    bsz = len(mols) + 3        # three empty segments at the end (short last batch)
    d["ref_gather_tanh"] = L.GraphGather(bsz, activation_fn=torch.tanh)(args).numpy()
    d["ref_gather_linear"] = L.GraphGather(bsz)(args).numpy()
    d["gather_batch_size"] = np.array(bsz)
    np.savez_compressed(os.path.join(HERE, "ref_layers.npz"), **d)
    print("layers: atoms", mm.get_num_atoms(), "conv", d["ref_conv_relu"].shape)

    # ---------------------------------------------------------------- model forward goldens
    for mode in ("classification", "regression"):
        pm = make_molecules(30, seed=31, shape="stress")
        mols = pm.to_list()
        cms, mm = ref_batch(mols)
        d = pack_mols(mols)
        d["features"] = d["features"].astype(np.float32)
        torch.manual_seed(4)
        n_tasks, bsz = 3, 32
        model = _GraphConvTorchModel(n_tasks, graph_conv_layers=[64, 64], number_input_features=[75, 64],
                                     dense_layer_size=128, dropout=0.0, mode=mode,
                                     number_atom_features=75, n_classes=2, batch_normalize=True,
                                     uncertainty=False, batch_size=bsz)
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() == 1:
                    p.add_(torch.randn_like(p) * 0.1)
        for k, v in model.state_dict().items():
            d["sd:" + k] = v.numpy().copy()
        args = layer_args(mm, len(mols))
        model.train()
        res = model(args)
        for i, r in enumerate(res):
            d["ref_train_out%d" % i] = r.detach().numpy()
        for k, v in model.state_dict().items():
            if "running" in k:
                d["sd_after:" + k] = v.numpy().copy()
        rng = np.random.default_rng(9)
        w = rng.random((len(mols), n_tasks)).astype(np.float32)
        if mode == "classification":
            yi = rng.integers(0, 2, size=(len(mols), n_tasks))
            y = np.eye(2, dtype=np.float32)[yi]
            crit = SoftmaxCrossEntropy()._create_pytorch_loss()
            losses = crit(res[1], torch.from_numpy(y))
        else:
            y = rng.standard_normal((len(mols), n_tasks)).astype(np.float32)
            crit = L2Loss()._create_pytorch_loss()
            losses = crit(res[0], torch.from_numpy(y))
        wt = torch.from_numpy(w)
        wt = wt.reshape(tuple(wt.shape) + (1,) * (losses.dim() - wt.dim()))
        d["y"], d["w"] = y, w
        d["ref_train_loss"] = (losses * wt).mean().detach().numpy()
        model.eval()
        res = model(args)
        for i, r in enumerate(res):
            d["ref_eval_out%d" % i] = r.detach().numpy()
        d["batch_size"] = np.array(bsz)
        np.savez_compressed(os.path.join(HERE, "ref_model_%s.npz" % mode), **d)
        print("model", mode, [r.shape for r in res], "loss", float(d["ref_train_loss"]))


if __name__ == "__main__":
    main()
