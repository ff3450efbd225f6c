import sys, time
sys.path.insert(0, "/root/repo")
import crick_b200 as cb
from crick_b200._lib import lib
row = 1000
for ng in (100000, 300000, 1000000):
    d = cb.DeviceArray((ng, row))
    lib.crick_synth_fill(d.ptr, None, ng * row, 2, 0, 2, None)
    lib.crick_device_synchronize()
    best = 1e9
    for rep in range(3):
        g = cb.TDigestGroup(ng, 100)
        lib.crick_device_synchronize()
        t0 = time.perf_counter(); g.update(d); g.flush(); lib.crick_device_synchronize()
        best = min(best, time.perf_counter() - t0)
    print("groups %7d  %8.3f ms (%5.2f Gval/s)" % (ng, best * 1e3, ng * row / best / 1e9), flush=True)
    del d
