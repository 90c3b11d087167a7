"""What does concurrent side-stream work cost the training kernels?  Resident training steps on the main stream
while a helper thread keeps a side stream busy with (a) nothing, (b) pinned H2D copies only, (c) the device
permute + record kernels only, (d) both — one side-stream job per training step (throttled by events)."""
import faulthandler
import itertools
import os
import sys
import threading
import time

faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepchem_b200 import ops
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules

dev = torch.device("cuda", 0)
B = 4096
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
m.model.train()
gen = m.default_generator(ds, epochs=1, deterministic=True, workers=1)
prepared = [m._prepare_batch(b) for b in itertools.islice(gen, 4)]
topo = prepared[0][0][1]._dcgc_topology
n = topo.n_atoms
host = torch.empty(n * 75, dtype=torch.float32, pin_memory=True).view(n, 75)
host.zero_()
fdev = torch.empty(n, 75, device=dev)
xout = torch.empty(n * 76, device=dev)
side = torch.cuda.Stream(device=dev)
K = 60


def run(mode, chunk_mb=0, d2h=False):
    stop = [False]
    tick = [torch.cuda.Event() for _ in range(K + 8)]

    def helper():
        torch.cuda.set_device(dev)
        with torch.cuda.stream(side):
            for i in range(K + 8):
                if stop[0]:
                    break
                while not tick[i].query():           # one job per training step
                    time.sleep(0.00005)
                    if stop[0]:
                        return
                if mode in ("h2d", "both"):
                    if chunk_mb:
                        flat_h, flat_d = host.view(-1), fdev.view(-1)
                        step = chunk_mb * (1 << 18)
                        for o in range(0, flat_h.numel(), step):
                            flat_d[o:o + step].copy_(flat_h[o:o + step], non_blocking=True)
                    else:
                        fdev.copy_(host, non_blocking=True)
                if mode in ("kernels", "both"):
                    ops.permute_rows(fdev, topo.perm, out=xout)
                if mode == "prepare":
                    m._prepare_batch(batches[i % len(batches)], slots[i % len(slots)])
            side.synchronize()

    th = threading.Thread(target=helper, daemon=True)
    for i in range(8):
        m._train_step(*prepared[i % 4])
    torch.cuda.synchronize()
    tick[0].record()
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = m._train_step(*prepared[i % 4])
        tick[i + 1].record()
        if d2h:
            loss_host[i % 4].copy_(loss.detach(), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    stop[0] = True
    th.join()
    print("side stream: %-8s chunk %2d MiB d2h %d env chunk %s -> %.3f ms/step" % (
        mode, chunk_mb, d2h, os.environ.get("DCGC_H2D_CHUNK_MB"), e0.elapsed_time(e1) / K))


from deepchem_b200.graphconvmodel import _DeviceSlot
gen2 = m.default_generator(ds, epochs=1000, deterministic=True, workers=2)
batches = [next(gen2) for _ in range(8)]
slots = [_DeviceSlot(dev) for _ in range(4)]
loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(4)]
with torch.cuda.stream(side):
    for i in range(8):
        m._prepare_batch(batches[i], slots[i % 4])
torch.cuda.synchronize()
run("none")
run("none", d2h=True)
os.environ["DCGC_H2D_CHUNK_MB"] = "0"
run("prepare")
os.environ["DCGC_H2D_CHUNK_MB"] = "1"
run("prepare")
run("prepare", d2h=True)
os.environ["DCGC_H2D_CHUNK_MB"] = "0.25"
run("prepare")
run("h2d", 1)
run("h2d")
run("none")
