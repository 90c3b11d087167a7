"""Isolate the tcgen05 wgrad discrepancy seen in the D-MPNN test (k1=147, a2=None, n=300)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepchem_b200 import ops, _lib
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
def run(rows, k, n, pad):
    ld = (k + 3) // 4 * 4 if pad else k
    buf = torch.randn(rows, ld, device=dev, generator=g)
    x = buf[:, :k]
    go = torch.randn(rows, n, device=dev, generator=g)
    ref = x.double().t() @ go.double()
    out = {}
    for name, mode in (("fp32", _lib.GEMM_FP32), ("tc", _lib.GEMM_TF32X3)):
        dw, db = ops.group_gemm_wgrad(x, None, go, None, 1, mode)
        err = (dw[0].double() - ref).abs()
        out[name] = float(err.max() / ref.abs().max())
        if name == "tc" and out[name] > 1e-4:
            bad = (err > 1e-3 * ref.abs().max()).nonzero()
            print("   bad entries:", bad.shape[0], "rows(k) range", int(bad[:, 0].min()), int(bad[:, 0].max()),
                  "cols(n) range", int(bad[:, 1].min()), int(bad[:, 1].max()))
        eb = float((db[0].double() - go.double().sum(0)).abs().max() / go.double().sum(0).abs().max())
        out[name + "_bias"] = eb
    print("rows=%d k=%d n=%d pad=%d:" % (rows, k, n, pad), {a: "%.2e" % b for a, b in out.items()})
for rows, k, n, pad in [(8000, 147, 300, 1), (8000, 148, 300, 0), (8000, 128, 300, 0), (8000, 147, 256, 1), (8000, 147, 128, 1),
                        (8000, 300, 300, 0), (8000, 433, 300, 0), (500, 300, 300, 0), (500, 300, 12, 0), (8000, 144, 128, 0)]:
    run(rows, k, n, pad)
