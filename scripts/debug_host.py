"""Debug helper: host layout-builder timings on the current box (pinned vs pageable output)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.synthetic import make_molecules
from deepchem_b200 import mol_graphs as MG, _lib
pm = make_molecules(4096, seed=0)
def t(fn, n=20):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print("cpus", os.cpu_count(), open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
print("build pageable ms", t(lambda: MG.BatchLayout.build(pm, n_segments=4096)))
if torch.cuda.is_available():
    st = torch.empty(8 << 20, dtype=torch.uint8, pin_memory=True)
    print("build into pinned staging ms", t(lambda: MG.BatchLayout.build(pm, n_segments=4096, staging=st)))
    print("build pinned alloc each time ms", t(lambda: MG.BatchLayout.build(pm, n_segments=4096, pinned=True)))
L = _lib.lib(); info = _lib.LayoutInfo()
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
print("plan only ms", t(lambda: L.dcgc_layout_plan(4096, P(pm.atom_ptr), P(pm.adj_ptr), 4096, 128, ctypes.byref(info))))
slab = np.empty(int(info.slab_bytes), np.uint8)
print("build only ms", t(lambda: L.dcgc_layout_build(4096, P(pm.atom_ptr), P(pm.adj_ptr), P(pm.adj_idx), ctypes.byref(info), P(slab))))
import threading
def worker(k):
    s = np.empty(int(info.slab_bytes), np.uint8)
    for _ in range(k): L.dcgc_layout_build(4096, P(pm.atom_ptr), P(pm.adj_ptr), P(pm.adj_idx), ctypes.byref(info), P(s))
for nt in (1, 2, 4):
    ths = [threading.Thread(target=worker, args=(20,)) for _ in range(nt)]
    t0 = time.perf_counter(); [x.start() for x in ths]; [x.join() for x in ths]
    print("threads", nt, "ms per batch (throughput)", (time.perf_counter() - t0) / (20 * nt) * 1e3)
