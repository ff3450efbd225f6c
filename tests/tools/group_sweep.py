"""Grouped reference-compatible ingest: warp-per-digest vs thread-per-digest by group count."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib

row = 1000
for ng in (1000, 2000, 5000, 10000, 20000, 50000, 100000, 300000):
    d = cb.DeviceArray((ng, row))
    lib.crick_synth_fill(d.ptr, None, ng * row, 2, 0, 2, None)
    lib.crick_device_synchronize()
    res = []
    for mode in (0, 1):
        lib.crick_group_kernel(mode)
        best = 1e9
        for rep in range(3):
            g = cb.TDigestGroup(ng, 100)
            lib.crick_device_synchronize()
            t0 = time.perf_counter(); g.update(d); g.flush(); lib.crick_device_synchronize()
            best = min(best, time.perf_counter() - t0)
        res.append(best)
    lib.crick_group_kernel(-1)
    print("groups %7d  warp/digest %8.3f ms (%5.2f Gval/s)   thread/digest %8.3f ms (%5.2f Gval/s)" % (
        ng, res[0] * 1e3, ng * row / res[0] / 1e9, res[1] * 1e3, ng * row / res[1] / 1e9), flush=True)
