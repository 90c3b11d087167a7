"""Per-role timeline of the tcgen05 wgrad kernel (CTA 0)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import _lib, mol_graphs as MG, ops
from deepchem_b200.synthetic import make_molecules
dev = torch.device("cuda", 0)
topo = MG.BatchLayout.build(make_molecules(4096, seed=0)).to_device(dev)
n = topo.n_atoms
x = torch.randn(n, 128, device=dev); s = torch.randn(n, 128, device=dev); g = torch.randn(n, 128, device=dev)
L = _lib.lib()
L.dcgcdbg_tc_timeline.argtypes = [ctypes.c_void_p]; L.dcgcdbg_tc_timeline.restype = None
for mode, name in ((_lib.GEMM_TF32X3, "tf32x3"), (_lib.GEMM_BF16, "bf16")):
    for _ in range(3):
        ops.group_gemm_wgrad(x, s, g, topo, 11, mode)
    buf = torch.zeros(8192, dtype=torch.int64, device=dev)
    L.dcgcdbg_tc_timeline(ctypes.c_void_p(buf.data_ptr()))
    ops.group_gemm_wgrad(x, s, g, topo, 11, mode)
    torch.cuda.synchronize()
    L.dcgcdbg_tc_timeline(None)
    t = buf.cpu().numpy().astype(np.int64)
    t0, steps = t[5000], int(t[5001])
    print("== %s: %d chunks of 32 atoms on CTA 0 (cycles from kernel start)" % (name, steps))
    print("producer commits:", (t[0:steps] - t0).tolist())
    print("mma saw chunk   :", (t[3072:3072 + steps] - t0).tolist())
    if t[1024] > 0:
        print("producer enters store:", (t[1024:1024 + steps] - t0).tolist())
        print("G loader commits     :", (t[2048:2048 + steps] - t0).tolist())
    print("epilogue: entered %d, accumulators ready %d, stores done %d" % (t[4096] - t0, t[4097] - t0, t[4098] - t0))
    gv = t[6000:6000 + 4 * 512].reshape(512, 4)
    gv = gv[gv[:, 0] > 0]
    if len(gv):
        st, en = gv[:, 0] - gv[:, 0].min(), gv[:, 1] - gv[:, 0].min()
        print("whole grid (%d CTAs): start ns min/median/max %d/%d/%d, end ns min/median/max %d/%d/%d; chunks min/max %d/%d" % (
            len(gv), st.min(), np.median(st), st.max(), en.min(), np.median(en), en.max(), gv[:, 2].min(), gv[:, 2].max()))
        dur = (gv[:, 1] - gv[:, 0]) / np.maximum(gv[:, 2], 1)
        print("  ns per chunk over CTAs: min %.0f median %.0f max %.0f; CTAs with < 20 chunks: %d" % (dur.min(), np.median(dur), dur.max(), int((gv[:, 2] < 20).sum())))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.group_gemm_wgrad(x, s, g, topo, 11, mode)
    e1.record(); torch.cuda.synchronize()
    print("wgrad (stage 1 + reduce) %.1f us" % (e0.elapsed_time(e1) * 100))
