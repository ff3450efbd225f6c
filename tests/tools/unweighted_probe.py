"""Throughput of update(x) without weights (8 B/value) next to update(x, w) (16 B/value)."""
import sys, time, ctypes
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
x = torch.empty(n, dtype=torch.float64, device="cuda")
w = torch.empty(n, dtype=torch.float64, device="cuda")
lib.crick_synth_fill(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()), n, 1234, 0, 0, None)
lib.crick_device_synchronize()
for name, args in (("x+w", (x, w)), ("x  ", (x,))):
    t = cb.TDigest(100)
    for _ in range(3):
        t.update(*args, mode="parallel", check_w=False)
    lib.crick_device_synchronize()
    lib.crick_profile_enable(1)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        t.update(*args, mode="parallel", check_w=False)
    lib.crick_device_synchronize()
    dt = (time.perf_counter() - t0) / reps
    ms = ctypes.c_double(); k = ctypes.c_int64()
    lib.crick_profile_collect(ctypes.byref(ms), ctypes.byref(k))
    lib.crick_profile_enable(0)
    b = 16 if len(args) == 2 else 8
    print("%s  step %.3f ms  %.1f Gval/s   kernel %.3f ms = %.0f GB/s (%.2f of 6542)" % (
        name, dt * 1e3, n / dt / 1e9, ms.value / k.value, b * n / (ms.value / k.value) / 1e6,
        b * n / (ms.value / k.value) / 1e6 / 6542.4), flush=True)
