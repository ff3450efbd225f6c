"""TDigest.update per-call time for small arrays (host NumPy input), auto vs forced modes."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib
rng = np.random.default_rng(0)
for n in (100, 1000, 2048, 4096, 8191, 8192, 20000):
    x = rng.lognormal(size=n)
    dx = cb.DeviceArray.from_numpy(x)
    out = []
    for mode in ("reference", "parallel"):
        if mode == "parallel" and n < 8192:
            out.append(float("nan")); continue
        t = cb.TDigest(100)
        t.update(dx, mode=mode); lib.crick_device_synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            t.update(dx, mode=mode)
        t.size(); lib.crick_device_synchronize()
        out.append((time.perf_counter() - t0) / 5 * 1e3)
    print("n=%6d  reference-mode %8.3f ms   parallel %8.3f ms" % (n, out[0], out[1]), flush=True)
