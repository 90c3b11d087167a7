"""Per-role timeline of the tcgen05 GEMM in its dgrad shape (K = 128, 256 output columns = two n-tiles), CTA (0, 0)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import _lib, mol_graphs as MG, ops
from deepchem_b200.synthetic import make_molecules
dev = torch.device("cuda", 0)
topo = MG.BatchLayout.build(make_molecules(4096, seed=0)).to_device(dev)
n = topo.n_atoms
g = torch.randn(n, 128, device=dev)
w = torch.randn(11, 256, 128, device=dev) / 16
L = _lib.lib()
L.dcgcdbg_tc_timeline.argtypes = [ctypes.c_void_p]; L.dcgcdbg_tc_timeline.restype = None
for _ in range(3):
    ops.group_gemm_dgrad(g, w, 128, 128, topo, True, True, _lib.GEMM_TF32X3)
buf = torch.zeros(8192, dtype=torch.int64, device=dev)
L.dcgcdbg_tc_timeline(ctypes.c_void_p(buf.data_ptr()))
ops.group_gemm_dgrad(g, w, 128, 128, topo, True, True, _lib.GEMM_TF32X3)
torch.cuda.synchronize()
L.dcgcdbg_tc_timeline(None)
t = buf.cpu().numpy().astype(np.int64)
t0 = t[5000]
total = 4
tiles = int((t[4096:4096 + 64:2] > 0).sum())
print("== dgrad tf32x3: %d tiles on CTA (0,0); cycles from kernel start" % tiles)
print("tile | mma first / last chunk seen | converter set 0 first/last | A TMA issued first/last | epilogue start / end")
for it in range(tiles):
    c0, c1 = it * total, it * total + total - 1
    g0 = t[0 + c0 // 2: 0 + c0 // 2 + total // 2] - t0
    print("%4d | %7d %7d | %7d %7d | %7d %7d | %7d %7d" % (it, t[3072 + c0] - t0, t[3072 + c1] - t0, g0[0], g0[-1],
                                                       t[2048 + c0] - t0, t[2048 + c1] - t0, t[4096 + 2 * it] - t0,
                                                       t[4096 + 2 * it + 1] - t0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.group_gemm_dgrad(g, w, 128, 128, topo, True, True, _lib.GEMM_TF32X3)
e1.record(); torch.cuda.synchronize()
print("dgrad %.1f us" % (e0.elapsed_time(e1) * 100))
