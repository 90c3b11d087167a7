"""Throughput of the MPNN message-passing phase (EdgeNetwork + GRU, T steps) and SetGather on Weave-shaped batches
(all n^2 pairs per molecule, 14 pair features, hidden 100: MPNNModel's defaults, graph_models.py:1063-1120) against
the CPU oracle on the box's host cores."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepchem_b200.mpnn import MessagePassing, SetGather
from oracle import mpnn_torch as M

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
B, P, h, F, T_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 14, 100, 75, 3
sizes = np.clip(rng.poisson(25, size=B), 6, 50)
a2p, start = [], 0
for n in sizes:
    C0, C1 = np.meshgrid(np.arange(n), np.arange(n))
    a2p.append(np.transpose(np.array([C1.flatten() + start, C0.flatten() + start])))
    start += n
a2p = np.concatenate(a2p).astype(np.int64)
n_atoms, n_pairs = int(sizes.sum()), a2p.shape[0]
pf = (rng.random((n_pairs, P)) < 0.25).astype(np.float32)
x = (rng.random((n_atoms, F)) < 0.1).astype(np.float32)
split = np.repeat(np.arange(B), sizes).astype(np.int32)
mp = MessagePassing(T_steps, n_hidden=h)
sg = SetGather(6, B, h)
pf_d, x_d, a2p_t = torch.from_numpy(pf).to(dev), torch.from_numpy(x).to(dev), torch.from_numpy(a2p)
for _ in range(3):
    out = mp([x_d, pf_d, a2p_t])
    q = sg([out, split])
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
K = 10
e0.record()
for _ in range(K):
    out = mp([x_d, pf_d, a2p_t])
e1.record()
for _ in range(K):
    q = sg([out, split])
e2.record()
torch.cuda.synchronize()
t_mp, t_sg = e0.elapsed_time(e1) / K, e1.elapsed_time(e2) / K
print("B=%d molecules, %d atoms, %d pairs, hidden %d, T=%d" % (B, n_atoms, n_pairs, h, T_steps))
print("GPU: message passing %.3f ms (%.0f molecules/s), SetGather(M=6) %.3f ms" % (t_mp, B / t_mp * 1e3, t_sg))
# CPU oracle on a bounded sample (the reference materialises an h x h matrix per pair: 40 KB per pair)
nb = min(B, 64)
na, npairs = int(sizes[:nb].sum()), int((sizes[:nb] ** 2).sum())
torch.set_num_threads(os.cpu_count() or 1)
enn, gru = mp.message_function, mp.update_function
args = (torch.from_numpy(x[:na]), torch.from_numpy(pf[:npairs]), torch.from_numpy(a2p[:npairs]), T_steps, h,
        (enn.W, enn.b), [getattr(gru, k) for k in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh")])
M.message_passing(*args)
t = time.perf_counter()
ref = M.message_passing(*args)
t_cpu = time.perf_counter() - t
print("CPU oracle (%d threads): message passing of %d molecules %.1f ms (%.0f molecules/s)" % (
    torch.get_num_threads(), nb, t_cpu * 1e3, nb / t_cpu))
err = float((out[:na].cpu() - ref).abs().max() / ref.abs().max())
print("max rel err vs fp32 oracle on the sample: %.2e" % err)

# ---- training: one MPNNModel step (forward + L2 loss + backward + fused Adam), reference defaults T = 5, M = 10
from deepchem_b200.mpnn import MPNNModel  # noqa: E402
torch.manual_seed(0)
Bt = min(B, 256)
na_t, np_t = int(sizes[:Bt].sum()), int((sizes[:Bt] ** 2).sum())
model = MPNNModel(12, n_atom_feat=F, n_pair_feat=P, n_hidden=h, T=5, M=10, batch_size=Bt)
inputs = [x[:na_t], pf[:np_t], split[:na_t], a2p[:np_t], Bt]
yb = rng.standard_normal((Bt, 12)).astype(np.float32)
wb = np.ones_like(yb)
gen = lambda n: (((inputs, [yb], [wb])) for _ in range(n))      # noqa: E731
model.fit_generator(gen(3))
torch.cuda.synchronize()
t = time.perf_counter()
model.fit_generator(gen(10))
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print("MPNNModel training step (T=5, M=10, %d molecules, %d pairs): %.2f ms = %.0f molecules/s (host loop included)" % (
    Bt, np_t, dt * 1e3, Bt / dt))
