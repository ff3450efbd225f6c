"""SpaceSaving(1000).update throughput on device-resident heavy-tailed int64 items (config 5
shape), counting path (mode 3) next to the shard path (mode 2)."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib

torch.manual_seed(3)
for n in (10**7, 10**8, 10**9):
    u = torch.rand(n, dtype=torch.float64, device="cuda")
    x = u.pow_(-10.0).clamp_(max=4.0e18).to(torch.int64)      # ~ Zipf(1.1): P(X >= k) = k^-0.1
    del u
    torch.cuda.synchronize()
    for mode in (3, 2):
        if mode == 2 and n > 10**8:
            continue
        best = 1e9
        for rep in range(2):
            s = cb.SpaceSaving(1000, "i8")
            lib.crick_device_synchronize()
            t0 = time.perf_counter()
            s.update(x, mode=mode)
            size = s.size()
            lib.crick_device_synchronize()
            best = min(best, time.perf_counter() - t0)
        c = s.counters()
        print("n=%.0e mode %d  %9.2f ms  %8.2f Gval/s  (%.0f GB/s)  kept %d  top %s  min count %d  max error %d" % (
            n, mode, best * 1e3, n / best / 1e9, 8 * n / best / 1e9, len(c), c["item"][:3].tolist(),
            c["count"].min(), c["error"].max()), flush=True)
    del x

# mid-size batches (the dask per-chunk shape): counting from 65536 elements
for n in (65536, 262144, 2 * 10**6):
    u = torch.rand(n, dtype=torch.float64, device="cuda")
    x = u.pow_(-10.0).clamp_(max=4.0e18).to(torch.int64)
    torch.cuda.synchronize()
    res = []
    for mode in (0, 2 if n >= 262144 else 1):
        s = cb.SpaceSaving(1000, "i8")
        lib.crick_device_synchronize()
        t0 = time.perf_counter(); s.update(x, mode=mode); s.size(); lib.crick_device_synchronize()
        res.append(time.perf_counter() - t0)
    print("n=%8d  auto (counting) %8.3f ms   %s %9.3f ms" % (n, res[0] * 1e3, "shards" if n >= 262144 else "replay", res[1] * 1e3), flush=True)

# small batches on a fresh summary, few distinct items (a categorical column)
for n in (8192, 30000, 60000):
    x = torch.randint(0, 500, (n,), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    res = []
    for mode in (0, 1):
        s = cb.SpaceSaving(1000, "i8")
        lib.crick_device_synchronize()
        t0 = time.perf_counter(); s.update(x, mode=mode); s.size(); lib.crick_device_synchronize()
        res.append(time.perf_counter() - t0)
    print("n=%6d 500 distinct  auto %7.3f ms   replay %8.3f ms" % (n, res[0] * 1e3, res[1] * 1e3), flush=True)
