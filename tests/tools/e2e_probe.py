"""bench.py's e2e leg in isolation: pinned host x + w (1e9 values each) -> TDigest.update -> centroids."""
import sys, time, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np
if len(sys.argv) > 2 and sys.argv[2] == "torch":
    import torch
    _s = torch.cuda.Stream(); _t = torch.zeros(1 << 20, device="cuda"); torch.cuda.synchronize()
import crick_b200 as cb
from crick_b200._lib import lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None)
lib.crick_device_synchronize()
hx, hw = cb.PinnedArray(n), cb.PinnedArray(n)
lib.crick_memcpy_d2h(hx.ptr, dx.ptr, 8 * n, None)
lib.crick_memcpy_d2h(hw.ptr, dw.ptr, 8 * n, None)
t = None
for rep in range(6):
    t0 = time.perf_counter()
    t = None                                   # the previous digest is freed inside the step, as in bench.py
    t1 = time.perf_counter()
    t = cb.TDigest(100, mode="parallel")
    t2 = time.perf_counter()
    t.update(hx.array, hw.array)
    t3 = time.perf_counter()
    c = t.centroids()
    dt = time.perf_counter() - t0
    print("step %d: %.1f ms  %.2f Gval/s  %.1f GB/s   (free %.1f, new %.1f, update %.1f, centroids %.1f ms)" % (
        rep, dt * 1e3, n / dt / 1e9, 16 * n / dt / 1e9, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (time.perf_counter() - t3) * 1e3), flush=True)
