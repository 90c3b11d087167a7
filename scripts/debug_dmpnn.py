import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import ops, _lib
from deepchem_b200.dmpnn import DMPNNModel, GraphDataset
from deepchem_b200.dmpnn_data import make_graphs
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
# --- raw GEMM shapes
for rows, k, n in [(8000, 300, 300), (8000, 147, 300), (8000, 128, 300), (8000, 300, 128), (500, 300, 300), (500, 303, 12)]:
    x = torch.randn(rows, k, device=dev, generator=g); w = torch.randn(k, n, device=dev, generator=g) / k ** 0.5
    b = torch.randn(n, device=dev, generator=g); go = torch.randn(rows, n, device=dev, generator=g)
    ref = x.double() @ w.double() + b.double()
    refd = go.double() @ w.double().t()
    for name, mode in (("fp32", _lib.GEMM_FP32), ("tc", _lib.GEMM_TF32X3)):
        y = ops.group_gemm_fwd(x, None, w, b, None, 0, mode)
        d1, _ = ops.group_gemm_dgrad(go, w, k, 0, None, True, False, mode)
        print("rows=%d k=%d n=%d %s: fwd %.2e dgrad %.2e" % (rows, k, n, name, rel(y, ref), rel(d1, refd)))
# two-operand (W_o shape)
rows = 4000
a1b = torch.zeros(rows, 136, device=dev); a1b[:, :133] = torch.randn(rows, 133, device=dev, generator=g); a1 = a1b[:, :133]
a2 = torch.randn(rows, 300, device=dev, generator=g); w = torch.randn(433, 300, device=dev, generator=g) / 20
go = torch.randn(rows, 300, device=dev, generator=g)
ref = torch.cat([a1, a2], 1).double() @ w.double(); refd = go.double() @ w.double().t()
for name, mode in (("fp32", _lib.GEMM_FP32), ("tc", _lib.GEMM_TF32X3)):
    y = ops.group_gemm_fwd(a1, a2, w, None, None, 0, mode)
    d1, d2 = ops.group_gemm_dgrad(go, w, 133, 300, None, False, True, mode)
    d1b, d2b = ops.group_gemm_dgrad(go, w, 133, 300, None, True, True, mode)
    print("W_o shape %s: fwd %.2e d2 %.2e d1(both) %.2e d2(both) %.2e" % (name, rel(y, ref), rel(d2, refd[:, 133:]), rel(d1b, refd[:, :133]), rel(d2b, refd[:, 133:])))
# --- model grads
pg = make_graphs(500, seed=3, shape="qm9", global_size=3, no_bond_fraction=0.03)
rng = np.random.default_rng(0)
y = rng.standard_normal((500, 12)).astype(np.float32); w = np.ones((500, 12), np.float32)
grads = {}
for mode in ("fp32", "tf32x3"):
    torch.manual_seed(0)
    m = DMPNNModel(device=dev, use_default_fdim=False, n_tasks=12, global_features_size=3, batch_size=500, gemm_mode=mode)
    batch = next(m.default_generator(GraphDataset(pg, y, w), deterministic=True))
    inputs, labels, weights = m._prepare_batch(batch)
    loss = m._loss(m.model(inputs), labels, weights); loss.backward()
    grads[mode] = {k: p.grad.clone() for k, p in m.model.named_parameters()}
for k in grads["fp32"]:
    print(k, "tc vs fp32 rel %.2e" % rel(grads["tf32x3"][k], grads["fp32"][k]))
