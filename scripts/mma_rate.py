#!/usr/bin/env python
"""Issue rate of tcgen05.mma (cycles per 128 x N x K instruction) in the forms the GEMM kernels use: operand A from shared
memory (SS) or tensor memory (TS), N 128 / 256, one or two accumulators, fixed or rotating operand tiles, kind::tf32 or
kind::f16.  python scripts/mma_rate.py > gpurun_out/mma_rate.md"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from deepchem_b200 import _lib  # noqa: E402

L = _lib.lib()
L.dcgcdbg_mma_rate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
L.dcgcdbg_mma_rate.restype = ctypes.c_int
dev = torch.device("cuda", 0)
out = torch.zeros(148, dtype=torch.int64, device=dev)
print("| kind | A from | N | accumulators | operand tiles | cycles / MMA (1 CTA) | cycles / MMA (148 CTAs, median) |")
print("|---|---|---:|---:|---|---:|---:|")
for v in (0, 1, 2, 3, 4, 5, 8, 9, 12, 13, 16, 17, 18, 19):
    row = []
    for ctas in (1, 148):
        best = None
        for reps in (256, 1280):
            for _ in range(3):
                _lib.check(L.dcgcdbg_mma_rate(v, reps, ctas, ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                torch.cuda.synchronize()
            t = out[:ctas].cpu().double().median().item()
            best = (reps, t) if best is None else best + (reps, t)
        r1, t1, r2, t2 = best
        row.append((t2 - t1) / (r2 - r1))       # slope: cycles per MMA without the fixed cost
    print("| %s | %s | %d | %d | %s | %.1f | %.1f |" % ("f16/bf16 K16" if v & 16 else "tf32 K8", "TMEM" if v & 1 else "smem",
                                                      256 if v & 2 else 128, 2 if v & 4 else 1,
                                                      "4 rotating" if v & 8 else "fixed", row[0], row[1]))

print()
print("| CTA pair (cta_group::2, M = 256) kind | A from | N | cycles / MMA (1 pair) | cycles / MMA (74 pairs, median) |")
print("|---|---|---:|---:|---:|")
for v in (32, 33, 34, 35):
    row = []
    for pairs in (1, 74):
        best = None
        for reps in (256, 1280):
            for _ in range(3):
                _lib.check(L.dcgcdbg_mma_rate(v, reps, pairs, ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                torch.cuda.synchronize()
            tt = out[:pairs].cpu().double().median().item()
            best = (reps, tt) if best is None else best + (reps, tt)
        r1, t1, r2, t2 = best
        row.append((t2 - t1) / (r2 - r1))
    print("| tf32 K8 | %s | %d | %.1f | %.1f |" % ("TMEM" if v & 1 else "smem", 256 if v & 2 else 128, row[0], row[1]))
