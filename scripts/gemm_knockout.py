"""Timing-only knockout study of the tcgen05 GEMM (the register-fed tcgen05 GEMM, tc_gemm_kernel_v4 with DCGC_TC_V4=1; written for its predecessor v3): which stage bounds it?
Each mask runs in its own process (DCGC_TC_KNOCKOUT is read once).  Results under a non-zero mask
are numerically wrong by construction; only the durations mean anything.

    python scripts/gemm_knockout.py            # driver: spawns one child per mask
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MASKS = [0, 1, 2, 4 | 16, 8, 32, 64, 128, 1 | 32, 31, 255]
NAMES = {1: "no-store", 2: "no-mma", 4: "no-Aload", 8: "no-Bcopy", 16: "no-Asts", 32: "no-tmem-ld", 64: "no-proxy-fence",
         128: "no-image-kernel"}


def child():
    import torch
    from deepchem_b200 import _lib, mol_graphs as MG, ops
    from deepchem_b200.synthetic import make_molecules
    dev = torch.device("cuda", 0)
    pm = make_molecules(4096, seed=0, shape="zinc")
    topo = MG.BatchLayout.build(pm).to_device(dev)
    n = topo.n_atoms
    x = torch.randn(n, 128, device=dev)
    s = torch.randn(n, 128, device=dev)
    w = torch.randn(11, 256, 128, device=dev) / 16
    b = torch.randn(11, 128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=10):
        ts = []
        for _ in range(3):
            fn()
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    t_fwd = timeit(lambda: ops.group_gemm_fwd(x, s, w, b, topo, 1, _lib.GEMM_TF32X3))
    t_dg = timeit(lambda: ops.group_gemm_dgrad(x, w, 128, 128, topo, True, True, _lib.GEMM_TF32X3))
    t_wg = timeit(lambda: ops.group_gemm_wgrad(x, s, x, topo, 11, _lib.GEMM_TF32X3))
    print("RESULT %s %.1f %.1f %.1f" % (os.environ.get("DCGC_TC_KNOCKOUT", "0"), t_fwd, t_dg, t_wg), flush=True)


def main():
    if os.environ.get("DCGC_KNOCKOUT_CHILD"):
        return child()
    print("| mask | knocked out | fwd K=256 N=128 us | dgrad K=128 N=256 us | wgrad us (not knocked) |")
    print("|---:|---|---:|---:|---:|")
    for m in MASKS:
        env = dict(os.environ, DCGC_TC_KNOCKOUT=str(m), DCGC_KNOCKOUT_CHILD="1")
        out = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True,
                             timeout=300)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        label = "+".join(v for k, v in NAMES.items() if m & k) or "-"
        if line:
            _, mm, a, b, c = line[0].split()
            print("| %s | %s | %s | %s | %s |" % (mm, label, a, b, c), flush=True)
        else:
            print("| %d | %s | FAILED | | |" % (m, label), flush=True)
            sys.stderr.write(out.stderr[-2000:])


if __name__ == "__main__":
    main()
