import numpy as np, time, sys, mmap
n = 1_280_000_000
mode = sys.argv[1]
t0 = time.perf_counter()
if mode == "plain":
    a = np.empty(n, dtype=np.uint8)
else:
    mm = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    mm.madvise(mmap.MADV_HUGEPAGE)
    a = np.frombuffer(mm, dtype=np.uint8)
a[::4096] = 1        # touch every page
t1 = time.perf_counter()
src = np.ones(16 << 20, dtype=np.uint8)
for o in range(0, n - (16 << 20), 16 << 20):
    np.copyto(a[o:o + (16 << 20)], src)
t2 = time.perf_counter()
print(mode, "touch %.3f s, copy %.3f s (%.1f GB/s)" % (t1 - t0, t2 - t1, n / (t2 - t1) / 1e9))
