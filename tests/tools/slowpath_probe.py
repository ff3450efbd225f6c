"""Throughput of the parallel update when no lookup grid resolves the edges (binary-search path)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crick_b200 as cb
from crick_b200._lib import lib
n = 10**8
rng = np.random.default_rng(0)
cases = {
    "lognormal": rng.lognormal(size=n),
    "bimodal narrow (+-1e6, sd 1e-3)": np.where(rng.random(n) < 0.5, 1e6, -1e6) + rng.normal(0, 1e-3, n),
    "normal + 1e-4 mass far outliers": np.where(rng.random(n) < 1e-4, 1e12, 0.0) + rng.normal(size=n),
    "integers 0..20 (ties)": rng.integers(0, 21, n).astype(float),
}
for name, x in cases.items():
    dx = cb.DeviceArray.from_numpy(x)
    t = cb.TDigest(100, mode="parallel")
    best = 1e9
    for _ in range(4):
        lib.crick_device_synchronize(); t0 = time.perf_counter()
        t.update(dx, 1)
        lib.crick_device_synchronize(); best = min(best, time.perf_counter() - t0)
    print("%-36s %7.1f Gval/s  centroids=%d" % (name, n / best / 1e9, len(t.centroids())), flush=True)
    del dx
