"""Per-call latency of TDigest.update on device-resident inputs, by chunk size (the dask
per-chunk shape).  Usage: [CRICK_CUDA_LIB=...] python tests/tools/latency_probe.py"""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib

torch.manual_seed(0)
for n in (10_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
    x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
    w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
    torch.cuda.synchronize()
    for weighted in (True, False):
        t = cb.TDigest(100)
        args = (x, w) if weighted else (x,)
        for _ in range(3):
            t.update(*args, mode="parallel", check_w=False)
        lib.crick_device_synchronize()
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            t.update(*args, mode="parallel", check_w=False)
        lib.crick_device_synchronize()
        dt = (time.perf_counter() - t0) / reps
        print("n=%10d %s  %8.1f us/call  %7.2f Gval/s  q50=%.6f" % (
            n, "x+w" if weighted else "x  ", dt * 1e6, n / dt / 1e9, float(t.quantile(0.5))), flush=True)
