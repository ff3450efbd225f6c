"""Throughput of the parallel update on other distributions / shapes (one GPU)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crick_b200 as cb
from crick_b200._lib import lib
n = int(float(os.environ.get("N", 5e8)))
dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
for kind, name in [(0, "lognormal+w"), (1, "uniform"), (2, "normal"), (3, "zipf ints")]:
    lib.crick_synth_fill(dx.ptr, dw.ptr, n, 7, 0, kind, None)
    lib.crick_device_synchronize()
    for weighted in (True, False):
        for comp in (100, 200, 1000):
            t = cb.TDigest(comp, mode="parallel")
            best = 1e9
            for _ in range(4):
                lib.crick_device_synchronize(); t0 = time.perf_counter()
                t.update(dx, dw if weighted else 1, check_w=False)
                lib.crick_device_synchronize(); best = min(best, time.perf_counter() - t0)
            bpv = 16 if weighted else 8
            print("%-12s w=%d c=%4d: %7.1f Gval/s  %6.0f GB/s  centroids=%d" % (
                name, weighted, comp, n / best / 1e9, bpv * n / best / 1e9, len(t.centroids())), flush=True)
