"""SummaryStats.update throughput on device-resident int64 / float64 arrays (8 B/value)."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib

for n in (10**8, 10**9):
    for dt_ in (torch.int64, torch.float64):
        if dt_ == torch.int64:
            x = torch.randint(1, 1 << 40, (n,), dtype=dt_, device="cuda")
        else:
            x = torch.randn(n, dtype=dt_, device="cuda")
        torch.cuda.synchronize()
        s = cb.SummaryStats()
        for _ in range(2):
            s.update(x)
        lib.crick_device_synchronize()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            s.update(x)
        lib.crick_device_synchronize()
        dt = (time.perf_counter() - t0) / reps
        print("n=%.0e %-8s %.3f ms  %.1f Gval/s  %.0f GB/s (%.2f of 6542)  mean=%.6g" % (
            n, str(dt_).split(".")[1], dt * 1e3, n / dt / 1e9, 8 * n / dt / 1e9, 8 * n / dt / 1e9 / 6542.4, s.mean()), flush=True)
        del x

# small batches: the sequential (bit-exact) kernel below 65536 values
for n in (1000, 10000, 65535, 65536, 200000):
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    s = cb.SummaryStats()
    s.update(x); lib.crick_device_synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        s.update(x)
    s.count(); lib.crick_device_synchronize()
    print("n=%7d  %.3f ms per update" % (n, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
