"""Per-role clock64() timeline of CTA 0 of a staged (molecule-group) kernel: where do producers / consumers wait?
    python scripts/mg_timeline.py [gather|gather_add|pool_fwd|pool_bwd]
kinds: 1 producer got the stage (empty acquired), 2 bulk copies issued + arrive,
       4 consumer warp 0 starts waiting on full, 5 full acquired, 6 group computed"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from deepchem_b200 import _lib, mol_graphs as MG  # noqa: E402
from deepchem_b200.engine import topology_struct  # noqa: E402
from deepchem_b200.synthetic import make_molecules  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "gather"
L = _lib.lib()
L.dcgcdbg_mg_timeline.argtypes = [ctypes.c_void_p]
L.dcgcdbg_mg_timeline.restype = None
if len(sys.argv) > 2:
    L.dcgcdbg_mg_mode(int(sys.argv[2]))   # knock-outs: 1 no stores, 2 no row reads, 3 no row copies
dev = torch.device("cuda", 0)
pm = make_molecules(4096, seed=0, shape="zinc")
topo = MG.BatchLayout.build(pm, n_segments=4096).to_device(dev)
N, W = topo.n_atoms, 128
x = torch.randn(N, W, device=dev)
add = torch.randn(N, W, device=dev)
out = torch.empty(N, W, device=dev)
arg = torch.randint(0, 4, (N, W), dtype=torch.uint8, device=dev)
ts = ctypes.byref(topology_struct(topo))
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731


def run():
    if what == "gather":
        _lib.check(L.dcgc_mg_gather_sum(p(x), W, ts, 0, W, None, 0, p(out), W, None))
    elif what == "gather_add":
        _lib.check(L.dcgc_mg_gather_sum(p(x), W, ts, 1, W, p(add), W, p(add), W, None))
    elif what == "pool_fwd":
        _lib.check(L.dcgc_mg_pool_fwd(p(x), W, None, None, ts, W, p(out), W, p(arg), W, None))
    else:
        _lib.check(L.dcgc_mg_pool_bwd(p(x), W, p(arg), W, None, ts, W, p(out), W, None))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record()
for _ in range(10):
    out.copy_(x)
c1.record()
torch.cuda.synchronize()
print("reference: torch copy_ of the same [N,%d] tensor %.2f us" % (W, c0.elapsed_time(c1) * 100))
print("%s: %.2f us per launch, n_groups %d max_rows %d max_entries %d" %
      (what, e0.elapsed_time(e1) * 100, topo.n_groups, topo.group_max_rows, topo.group_max_entries))
buf = torch.zeros(8 * 500, dtype=torch.int64, device=dev)
L.dcgcdbg_mg_timeline(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
L.dcgcdbg_mg_timeline(None)
h = buf.cpu().numpy().reshape(500, 8)
t0 = h[h > 0].min()
print("it | P acquired  copies issued  - | C wait  acquired  done   (cycles from the first mark)")
for it in range(500):
    if not h[it].any():
        break
    print("%3d | %8d %8d %8d | %8d %8d %8d" % ((it,) + tuple(int(h[it, k] - t0) if h[it, k] else -1 for k in range(1, 7))))
