"""Per-role timeline of one tcgen05 GEMM (CTA 0): where does a tile's time go?"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import _lib, mol_graphs as MG, ops
from deepchem_b200.synthetic import make_molecules
dev = torch.device("cuda", 0)
topo = MG.BatchLayout.build(make_molecules(4096, seed=0)).to_device(dev)
n = topo.n_atoms
x = torch.randn(n, 128, device=dev); s = torch.randn(n, 128, device=dev)
w = torch.randn(11, 256, 128, device=dev) / 16; b = torch.randn(11, 128, device=dev)
L = _lib.lib()
L.dcgcdbg_tc_timeline.argtypes = [ctypes.c_void_p]; L.dcgcdbg_tc_timeline.restype = None
for mode, name in ((_lib.GEMM_TF32X3, "tf32x3"), (_lib.GEMM_F16X3, "f16x3 (64-wide chunks: 4 per tile)"), (_lib.GEMM_BF16, "bf16")):
    for _ in range(3):
        ops.group_gemm_fwd(x, s, w, b, topo, 1, mode)
    buf = torch.zeros(6000, dtype=torch.int64, device=dev)
    L.dcgcdbg_tc_timeline(ctypes.c_void_p(buf.data_ptr()))
    ops.group_gemm_fwd(x, s, w, b, topo, 1, mode)
    torch.cuda.synchronize()
    L.dcgcdbg_tc_timeline(None)
    t = buf.cpu().numpy().astype(np.int64)
    t0 = t[5000]
    total = 4 if mode == _lib.GEMM_F16X3 else 8
    tiles = int((t[4096:4096 + 64:2] > 0).sum())
    print("== %s: %d tiles on CTA 0; cycles relative to kernel start (1 cycle = 0.51 ns at 1965 MHz)" % (name, tiles))
    print("tile | mma first chunk seen | mma last chunk seen | prod0 first/last commit | prod1 first/last | B first/last issue | epi start | epi end")
    for it in range(tiles):
        c0, c1 = it * total, it * total + total - 1
        g0 = t[0 + c0 // 2: 0 + c0 // 2 + total // 2] - t0
        g1 = t[1024 + c0 // 2: 1024 + c0 // 2 + total // 2] - t0
        print("%4d | %7d | %7d | %7d %7d | %7d %7d | %7d %7d | %7d | %7d" % (
            it, t[3072 + c0] - t0, t[3072 + c1] - t0, g0[0], g0[-1], g1[0], g1[-1], t[2048 + c0] - t0, t[2048 + c1] - t0,
            t[4096 + 2 * it] - t0, t[4096 + 2 * it + 1] - t0))
    e = (t[5100:5116] - t[5100]).reshape(4, 4)
    print("epilogue of tile 2, warp 0, per 32-column block: [start, after tcgen05.ld, after STS+syncwarp, after stores]")
    print(e.tolist())
    pr = (t[5200:5200 + 4 * 24] - t0).reshape(24, 4)
    print("producer warp 8 (set 0), per chunk [enter commit, stage free, stores issued, stores done]:")
    print(pr[8:20].tolist())
    mm = t[3072:3072 + tiles * total] - t0
    print("mma chunk-to-chunk deltas (cycles):", np.diff(mm)[:24].tolist())
