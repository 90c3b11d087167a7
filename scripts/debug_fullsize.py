"""Debug helper: per-parameter gradient error of the CUDA model vs the fp32 and fp64 oracle at B=4096."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import oracle_batch, torch_args
from oracle import graphconv_torch as O
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import make_labels, make_molecules

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pm = make_molecules(B, seed=1)
y, w = make_labels(B, 1, "regression", seed=3)
torch.manual_seed(7)
layers = [128, 128, 128]
om = O.OracleGraphConvModel(1, layers, 128, mode="regression", batch_size=B)
m = GraphConvModel(1, layers, 128, mode="regression", batch_size=B)
m.model.load_state_dict(om.state_dict())
batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
inputs, labels, weights = m._prepare_batch(batch)
m.model.train()
outs = m.model(inputs)
loss = m._loss_fn([outs[0]], labels, weights)
loss.backward()
_, mm = oracle_batch(pm.to_list())
res = {}
for dt in (torch.float32, torch.float64):
    o2 = O.OracleGraphConvModel(1, layers, 128, mode="regression", batch_size=B).to(dt)
    o2.load_state_dict({k: v.to(dt) if v.is_floating_point() else v for k, v in om.state_dict().items()})
    o2.train()
    oo = o2(torch_args(mm, B, dtype=dt))
    lo = O.standard_loss("regression", oo, torch.from_numpy(y).to(dt), torch.from_numpy(w).to(dt))
    lo.backward()
    res[dt] = (oo, lo, dict(o2.named_parameters()))
o32, o64 = res[torch.float32], res[torch.float64]
print("loss ours %.8f o32 %.8f o64 %.8f" % (float(loss), float(o32[1]), float(o64[1])))
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))
print("out: ours-o64 %.2e o32-o64 %.2e ours-o32 %.2e" % (rel(outs[0].detach().cpu(), o64[0][0].detach()), rel(o32[0][0].detach(), o64[0][0].detach()), rel(outs[0].detach().cpu(), o32[0][0].detach())))
rows = []
for name, p in m.model.named_parameters():
    g64 = o64[2][name].grad
    g32 = o32[2][name].grad
    if g64 is None or p.grad is None:
        continue
    if float(g64.abs().max()) == 0:
        continue
    rows.append((rel(p.grad.cpu(), g64), rel(g32, g64), float(g64.abs().max()), name))
rows.sort(reverse=True)
for r in rows[:25]:
    print("ours-o64 %.2e  o32-o64 %.2e  |g|max %.2e  %s" % r)
