"""Config 4's merge: 65536 digests (compression 200, 2000 values each) into one, by the exact
pairwise tree (merge_tree) and by TDigestGroup.merge_all() (one parallel k-bucket update over all
centroids).  Run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import json, sys, time
sys.path.insert(0, "/root/repo")
import crick_b200 as cb
from crick_b200._lib import lib

nd, per = 65536, 2000
d = cb.DeviceArray((nd, per))
lib.crick_synth_fill(d.ptr, None, nd * per, 3, 0, 2, None)
g = cb.TDigestGroup(nd, 200)
g.update(d); g.flush(); lib.crick_device_synchronize()
out = {}
for name in ("merge_all", "merge_tree"):
    best = 1e9
    for rep in range(3):
        lib.crick_device_synchronize(); t0 = time.perf_counter()
        m = getattr(g, name)()
        lib.crick_device_synchronize(); best = min(best, time.perf_counter() - t0)
    out[name] = {"ms": best * 1e3, "centroids": int(len(m.centroids()))}
print(json.dumps(out))
