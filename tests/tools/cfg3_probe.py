"""Config 3 (1e6 digests x 1000 values, reference mode) timing for a given decision margin.
usage: cfg3_probe.py [ngroups] [margin]"""
import ctypes, json, sys, time
import numpy as np
sys.path.insert(0, ".")
import crick_b200 as cb
from crick_b200._lib import lib

ng = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6
margin = float(sys.argv[2]) if len(sys.argv) > 2 else 6e-13
row = 1000
lib.crick_group_margin(margin)
d = cb.DeviceArray((ng, row))
lib.crick_synth_fill(d.ptr, None, ng * row, 2, 0, 2, None)
lib.crick_device_synchronize()
nfb = ctypes.c_int64(0)
ts = []
for it in range(4):
    lib.crick_group_exact_fallbacks(ctypes.byref(nfb), 1)
    t0 = time.perf_counter()
    g = cb.TDigestGroup(ng, 100)
    g.update(d)
    g.flush()
    lib.crick_device_synchronize()
    ts.append(time.perf_counter() - t0)
    lib.crick_group_exact_fallbacks(ctypes.byref(nfb), 1)
t = min(ts[1:])
print(json.dumps({"ngroups": ng, "margin": margin, "ms": t * 1e3, "gvalues_s": ng * row / t / 1e9,
                  "fallbacks": nfb.value, "centroids0": int(len(g.centroids(0)))}))
