"""End-to-end cost of TDigest.update on host arrays (pageable vs pinned), split by phase."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crick_b200 as cb
from crick_b200._lib import lib
for n in (10**7, 10**8, 4 * 10**8):
    rng = np.random.default_rng(0)
    x = rng.lognormal(size=n); w = rng.uniform(0.5, 1.5, n)
    px, pw = cb.PinnedArray(n), cb.PinnedArray(n)
    px.array[:] = x; pw.array[:] = w
    for name, (a, b) in {"pageable": (x, w), "pinned": (px.array, pw.array)}.items():
        for _ in range(3):
            t0 = time.perf_counter(); t = cb.TDigest(100, mode="parallel"); t1 = time.perf_counter()
            t.update(a, b); t2 = time.perf_counter(); c = t.centroids(); t3 = time.perf_counter()
        print("n=%.0e %-9s new %.4f  update %.4f  centroids %.4f  -> %.2f Gval/s  %.1f GB/s" % (
            n, name, t1 - t0, t2 - t1, t3 - t2, n / (t3 - t0) / 1e9, 16 * n / (t3 - t0) / 1e9), flush=True)
    del px, pw
