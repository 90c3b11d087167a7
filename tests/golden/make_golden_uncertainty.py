"""Uncertainty-mode fixture from the reference checkout.  Run in the build container:

    python tests/golden/make_golden_uncertainty.py

  ref_model_uncertainty.npz   30 'stress' molecules through the REFERENCE's ConvMol / agglomerate_mols and its
        GraphConvModel(mode='regression', uncertainty=True, dropout=0.25, [64, 64], dense 128): the five outputs
        [y, exp(log_var), y, log_var, fingerprint] (graphconvmodel.py:238-246) in train mode (BatchNorm batch statistics;
        the module's own ``training`` argument stays False, so no dropout mask is drawn) and in eval mode,
        and the reference's uncertainty loss closure (graphconvmodel.py:360-372) on the 'loss' outputs with random
        labels and weights, called through the reference's own ``_loss_fn``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, HERE)
    import make_golden as G            # imports the reference (rdkit stubbed) and the helpers
    import torch
    from deepchem.models.torch_models.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_molecules
    pm = make_molecules(30, seed=41, shape="stress")
    mols = pm.to_list()
    cms, mm = G.ref_batch(mols)
    d = G.pack_mols(mols)
    d["features"] = d["features"].astype(np.float32)
    n_tasks, bsz = 3, 32
    torch.manual_seed(14)
    ref = GraphConvModel(n_tasks, number_input_features=[75, 64], graph_conv_layers=[64, 64], dense_layer_size=128,
                         dropout=0.25, mode="regression", uncertainty=True, batch_size=bsz, device=torch.device("cpu"))
    assert ref.output_types == ['prediction', 'variance', 'loss', 'loss', 'embedding']
    model = ref.model
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    for k, v in model.state_dict().items():
        d["sd:" + k] = v.numpy().copy()
    args = G.layer_args(mm, len(mols))
    model.train()
    res = model(args)
    for i, r in enumerate(res):
        d["ref_train_out%d" % i] = r.detach().numpy()
    rng = np.random.default_rng(19)
    y = rng.standard_normal((len(mols), n_tasks)).astype(np.float32)
    w = rng.random((len(mols), n_tasks)).astype(np.float32)
    loss = ref._loss_fn([res[i] for i in ref._loss_outputs], [torch.from_numpy(y)], [torch.from_numpy(w)])
    d["y"], d["w"] = y, w
    d["ref_train_loss"] = loss.detach().numpy()
    model.eval()
    for i, r in enumerate(model(args)):
        d["ref_eval_out%d" % i] = r.detach().numpy()
    d["batch_size"] = np.array(bsz)
    np.savez_compressed(os.path.join(HERE, "ref_model_uncertainty.npz"), **d)
    print("uncertainty model:", [tuple(r.shape) for r in res], "loss", float(loss))


if __name__ == "__main__":
    main()
