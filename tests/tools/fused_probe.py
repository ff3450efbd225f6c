"""TDigest.update(x, w, stats=SummaryStats) -- both sketches in one pass -- vs the two passes."""
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

def timed(fn, reps=8):
    for _ in range(2): fn()
    lib.crick_device_synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    lib.crick_device_synchronize()
    return (time.perf_counter() - t0) / reps

t = cb.TDigest(100); s = cb.SummaryStats()
fused = timed(lambda: t.update(x, w, mode="parallel", check_w=False, stats=s))
t2 = cb.TDigest(100); s2 = cb.SummaryStats()
two = timed(lambda: (t2.update(x, w, mode="parallel", check_w=False), s2.update(x)))
print("fused %.3f ms (%.1f Gval/s)   separate %.3f ms (%.1f Gval/s)   mean %.9g / %.9g" % (
    fused * 1e3, n / fused / 1e9, two * 1e3, n / two / 1e9, s.mean(), s2.mean()))
