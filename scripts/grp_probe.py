import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib
ng, row = 200000, 1000
d = cb.DeviceArray((ng, row))
lib.crick_synth_fill(d.ptr, None, ng * row, 2, 0, 2, None)
lib.crick_device_synchronize()
for _ in range(2):
    g = cb.TDigestGroup(ng, 100)
    t0 = time.perf_counter(); g.update(d); g.flush(); lib.crick_device_synchronize(); print(time.perf_counter() - t0)
