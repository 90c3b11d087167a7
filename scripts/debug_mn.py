import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepchem_b200 import ops, _lib
dev = torch.device("cuda", 0)
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
rows, k, n = 32, 128, 128
f = torch.arange(1, k + 1, device=dev, dtype=torch.float32)
x = f.repeat(rows, 1).contiguous()
g = torch.zeros(rows, n, device=dev); g[:, 0] = 1
dw, db = ops.group_gemm_wgrad(x, None, g, None, 1, _lib.GEMM_BF16)
print("test1 dW[:,0]/32 (expect 1..128):", (dw[0][:, 0] / 32)[:40].tolist())
print("test1 nonzero columns:", (dw[0].abs().sum(0) > 0).nonzero().flatten().tolist()[:20])
g = torch.zeros(rows, n, device=dev); g[:, 5] = 1
dw, db = ops.group_gemm_wgrad(x, None, g, None, 1, _lib.GEMM_BF16)
print("test2 nonzero columns (expect [5]):", (dw[0].abs().sum(0) > 0).nonzero().flatten().tolist()[:20])
# K consistency: x row r0 only, g row r1 only
for r0, r1 in ((0, 0), (3, 3), (3, 4), (9, 9), (17, 17), (31, 31), (8, 0)):
    x = torch.zeros(rows, k, device=dev); x[r0] = f
    g = torch.zeros(rows, n, device=dev); g[r1, 0] = 1
    dw, _ = ops.group_gemm_wgrad(x, None, g, None, 1, _lib.GEMM_BF16)
    print("rows (%d,%d): dW[:8,0] =" % (r0, r1), dw[0][:8, 0].tolist(), "sum", float(dw[0].abs().sum()))
