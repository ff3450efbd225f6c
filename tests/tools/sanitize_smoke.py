"""Small inputs through every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib

rng = np.random.default_rng(0)
# parallel t-digest update (sample sort, run merge, edges, accumulate, finalize), weighted + not
x = rng.lognormal(size=70001); w = rng.uniform(0.5, 1.5, x.size)
t = cb.TDigest(100)
t.update(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w), mode="parallel")
t.update(cb.DeviceArray.from_numpy(x), mode="parallel")
s = cb.SummaryStats()
t.update(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w), mode="parallel", stats=s)
print("tdigest", t.size(), len(t.centroids()), t.quantile(0.5), s.mean())
t2 = cb.TDigest(333.5); t2.update(x, mode="parallel"); print("c=333.5", len(t2.centroids()))
# reference mode, merges, bulk record merge
r = cb.TDigest(100, mode="reference"); r.update(x[:5000]); r.merge(t); print("ref", r.size())
# grouped: both kernels
for mode in (0, 1):
    lib.crick_group_kernel(mode)
    g = cb.TDigestGroup(70, 100)
    g.update(rng.normal(size=(70, 300)))
    g.update(rng.normal(size=(70, 300)), rng.uniform(1, 2, (70, 300)))
    print("group kernel", mode, g.meta()[:2, 0], g.quantile([0.5])[:2, 0])
lib.crick_group_kernel(-1)
# stats block path
st = cb.SummaryStats(); st.update(cb.DeviceArray.from_numpy(rng.normal(size=300001))); print("stats", st.count(), st.var())
# SpaceSaving: replay, shards, counting (exact + approximate), records
it = (rng.zipf(1.2, 60000) % 500).astype(np.int64)
for mode in (1, 2, 3):
    sp = cb.SpaceSaving(100, "i8"); sp.update(cb.DeviceArray.from_numpy(it), mode=mode)
    print("spsv mode", mode, sp.size(), sp.counters()[:2].tolist())
sp2 = cb.SpaceSaving(1000, "i8"); sp2.update(cb.DeviceArray.from_numpy(it), mode=3); print("spsv exact", sp2.size())
buf = cb.DeviceArray((2, sp.record_int64s()), "i8")
sp.export_record(buf.ptr); sp.export_record(buf.ptr + 8 * sp.record_int64s())
o = cb.SpaceSaving(100, "i8"); o.merge_records(buf.ptr, 2); print("records", o.size())
