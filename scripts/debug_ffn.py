import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import ops, _lib
from deepchem_b200._lib import ACT_NONE, ACT_RELU
dev = torch.device("cuda", 0)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))
torch.manual_seed(0)
x = torch.randn(500, 303, device=dev)
W0 = (torch.randn(303, 300, device=dev) / 17).requires_grad_(True); b0 = torch.randn(300, device=dev).requires_grad_(True)
W1 = (torch.randn(300, 300, device=dev) / 17).requires_grad_(True); b1 = torch.randn(300, device=dev).requires_grad_(True)
W2 = (torch.randn(300, 12, device=dev) / 17).requires_grad_(True); b2 = torch.randn(12, device=dev).requires_grad_(True)
dout = torch.randn(500, 12, device=dev) * 1e-4
def run(mode):
    for p in (W0, b0, W1, b1, W2, b2): p.grad = None
    inter = {}
    h0 = ops.GroupLinear2Fn.apply(x, None, W0, b0, ACT_RELU, mode); h0.retain_grad()
    h1 = ops.GroupLinear2Fn.apply(h0, None, W1, b1, ACT_RELU, mode); h1.retain_grad()
    o = ops.GroupLinear2Fn.apply(h1, None, W2, b2, ACT_NONE, mode)
    o.backward(dout)
    return dict(h0=h0.detach(), h1=h1.detach(), o=o.detach(), dh1=h1.grad, dh0=h0.grad, dW0=W0.grad, dW1=W1.grad, dW2=W2.grad,
                db0=b0.grad, db1=b1.grad, db2=b2.grad)
a = run(_lib.GEMM_FP32); b = run(_lib.GEMM_TF32X3)
for k in a: print(k, "%.2e" % rel(b[k], a[k]))
# small dout scale check of the dgrad alone
for scale in (1.0, 1e-4, 1e-8):
    g = torch.randn(500, 12, device=dev) * scale
    w = W2.detach().contiguous()
    d32, _ = ops.group_gemm_dgrad(g, w, 300, 0, None, True, False, _lib.GEMM_FP32)
    dtc, _ = ops.group_gemm_dgrad(g, w, 300, 0, None, True, False, _lib.GEMM_TF32X3)
    print("dgrad K=12 scale", scale, "%.2e" % rel(dtc, d32))
