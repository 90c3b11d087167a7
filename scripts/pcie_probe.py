"""Host <-> device copy bandwidth of one process (pinned memory), to see what two ranks of one box share.
python scripts/pcie_probe.py [seconds]"""
import sys, time, torch
dev = torch.device("cuda", 0)
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < secs:
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); k += 1
    dt = time.perf_counter() - t0
    print("%s: %.1f GB/s per direction" % (name, k * n / dt / 1e9), flush=True)
