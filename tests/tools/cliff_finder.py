"""Per-call time of every public entry point over a log grid of sizes: look for cliffs."""
import sys, time, pickle
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib
rng = np.random.default_rng(0)

def timed(fn, reps=3):
    fn(); lib.crick_device_synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    lib.crick_device_synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

sizes = [10, 100, 1000, 2047, 2048, 10**4, 10**5, 10**6, 10**7]
print("TDigest.update (NumPy host / device array), SummaryStats.update, SpaceSaving.update [ms]")
for n in sizes:
    x = rng.lognormal(size=n); dx = cb.DeviceArray.from_numpy(x)
    it = (rng.zipf(1.2, n) % 5000).astype(np.int64); dit = cb.DeviceArray.from_numpy(it)
    t = cb.TDigest(100); st = cb.SummaryStats(); sp = cb.SpaceSaving(1000, "i8")
    row = [timed(lambda: t.update(x)), timed(lambda: t.update(dx)), timed(lambda: t.update(dx, 2.0)),
           timed(lambda: st.update(x)), timed(lambda: st.update(dx)),
           timed(lambda: sp.update(it)), timed(lambda: sp.update(dit))]
    print("n=%9d  td host %8.3f dev %8.3f dev,w=2 %8.3f | stats host %8.3f dev %8.3f | spsv host %8.3f dev %8.3f" % ((n,) + tuple(row)), flush=True)
print("queries on a 1e6-value digest [ms]")
t = cb.TDigest(100); t.update(rng.normal(size=10**6))
for nq in (1, 100, 10**4, 10**6, 10**7):
    q = rng.uniform(size=nq); dq = cb.DeviceArray.from_numpy(q)
    print("nq=%9d  quantile host %8.3f dev %8.3f   cdf host %8.3f dev %8.3f" % (
        nq, timed(lambda: t.quantile(q)), timed(lambda: t.quantile(dq)), timed(lambda: t.cdf(q)), timed(lambda: t.cdf(dq))), flush=True)
print("merge / pickle / grouped [ms]")
parts = []
for k in range(64):
    p = cb.TDigest(100); p.update(rng.normal(size=20000)); parts.append(p)
m = cb.TDigest(100)
print("merge 64 digests %8.3f   pickle round trip %8.3f   centroids %8.3f" % (
    timed(lambda: cb.TDigest(100).merge(*parts)), timed(lambda: pickle.loads(pickle.dumps(parts[0]))), timed(lambda: parts[0].centroids())))
for ng, row in ((10, 1000), (1000, 1000), (20000, 200), (100000, 50)):
    v = rng.normal(size=(ng, row))
    dv = cb.DeviceArray.from_numpy(v)
    print("grouped %7d x %5d   host %9.3f   dev %9.3f" % (ng, row, timed(lambda: cb.grouped_update(v)), timed(lambda: cb.grouped_update(dv))), flush=True)
