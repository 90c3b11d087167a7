"""Layer-by-layer comparison of the CUDA per-layer path against the CPU oracle on the real Delaney molecules."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import torch.nn.functional as F
from helpers import GOLDEN, load_golden, oracle_batch, rel_err, torch_args, unpack_mols
from deepchem_b200.data import CSVLoader, PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200 import ops
from deepchem_b200._lib import ACT_RELU
from oracle import graphconv_torch as O

d = load_golden("ref_delaney_real.npz")
n = int(d["batch_size"])
which = sys.argv[1] if len(sys.argv) > 1 else "csv"
if which == "csv":
    ds = CSVLoader(["y"]).create_dataset(os.path.join(GOLDEN, "delaney.csv")).select_range(0, n)
else:
    from deepchem_b200.synthetic import PackedMols
    ds = PackedDataset(PackedMols(d["atom_ptr"], d["adj_ptr"], d["adj_idx"], d["features"]), d["y"], d["w"])
sd = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd:")}
mols = unpack_mols(d)
_, mm = oracle_batch(mols)
om = O.OracleGraphConvModel(1, [64, 64], 128, mode="regression", batch_size=n)
om.load_state_dict(sd)
om.eval()
args = torch_args(mm, n)
print("oracle vs golden eval:", rel_err(om(args)[0].detach().numpy(), d["ref_eval_out0"]))
for mode in ("fp32", "tf32x3"):
    m = GraphConvModel(1, [64, 64], 128, mode="regression", batch_size=n, gemm_mode=mode)
    m.model.load_state_dict(sd)
    for gen_mode in ("fit", "predict"):
        batch = next(m.default_generator(ds, mode=gen_mode, deterministic=True, pad_batches=False))
        inputs, labels, weights = m._prepare_batch(batch)
        print(mode, gen_mode, "inputs:", [tuple(t.shape) for t in inputs[:4]], "n_samples", inputs[3])
        print("  features vs oracle order:", rel_err(inputs[0].cpu().numpy()[:, :75], args[0].numpy()),
              "deg_slice eq", np.array_equal(inputs[1].cpu().numpy(), args[1].numpy()),
              "membership eq", np.array_equal(inputs[2].cpu().numpy(), args[2].numpy()))
        m.model.eval()
        mod = m.model
        with torch.no_grad():
            h, ho = inputs[0], args[0]
            adjs_o = [a.long() for a in args[4:]]
            for i in range(2):
                h = mod.graph_convs[i]([h, inputs[1], inputs[2]] + list(inputs[4:]))
                ho = om.graph_convs[i]([ho, args[1], args[2].long()] + adjs_o)
                print("  conv%d" % i, rel_err(h.cpu().numpy(), ho.numpy()))
                h = mod.batch_norms[i](h); ho = om.batch_norms[i](ho)
                print("  bn%d" % i, rel_err(h.cpu().numpy(), ho.numpy()))
                h = mod.graph_pools[i]([h, inputs[1], inputs[2]] + list(inputs[4:]))
                ho = O.graph_pool(ho, args[1], adjs_o)
                print("  pool%d" % i, rel_err(h.cpu().numpy(), ho.numpy()))
            h = ops.GroupLinearFn.apply(h, mod.dense.weight.t().contiguous(), mod.dense.bias, ACT_RELU, mod.gemm_mode)
            ho = F.relu(om.dense(ho))
            print("  dense", rel_err(h.cpu().numpy(), ho.numpy()))
            h = mod.batch_norms[-1](h); ho = om.batch_norms[-1](ho)
            h = mod.graph_gather([h, inputs[1], inputs[2]] + list(inputs[4:]))
            ho = O.graph_gather(ho, args[2].long(), n, torch.tanh)
            print("  gather", rel_err(h.cpu().numpy(), ho.numpy()))
            out = m.model(inputs)
            print("  model eval out vs golden:", rel_err(out[0].cpu().numpy(), d["ref_eval_out0"]))
