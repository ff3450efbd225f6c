import sys, time
sys.path.insert(0, "/root/repo")
import crick_b200 as cb
from crick_b200._lib import lib
nd, per = 65536, 2000
d = cb.DeviceArray((nd, per))
lib.crick_synth_fill(d.ptr, None, nd * per, 3, 0, 2, None)
lib.crick_device_synchronize()
for rep in range(3):
    g2 = cb.TDigestGroup(nd, 200)
    g2.update(d); g2.flush(); lib.crick_device_synchronize()
    o = cb.TDigest(200); lib.crick_device_synchronize()
    t0 = time.perf_counter()
    from crick_b200._lib import check
    check(lib.crick_tdigest_group_merge_tree(g2._h, o._h), "mt"); lib.crick_device_synchronize()
    t1 = time.perf_counter()
    g3 = cb.TDigestGroup(nd, 200); g3.update(d); g3.flush(); lib.crick_device_synchronize()
    t2 = time.perf_counter(); m = g3.merge_tree(); lib.crick_device_synchronize(); t3 = time.perf_counter()
    print("C call %.2f ms   python merge_tree() %.2f ms" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3), flush=True)
