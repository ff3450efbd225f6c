"""First-contact GPU script: exercises each path once and prints PASS/FAIL lines."""
import hashlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crick_b200 as cb
from crick_b200 import TDigest, SummaryStats, SpaceSaving, TDigestGroup, DeviceArray
from crick_b200._lib import lib
import oracle

G = json.load(open("tests/golden/golden.json"))
def sha(t): return hashlib.sha256(t.centroids().tobytes()).hexdigest()
def ok(name, cond, extra=""):
    print(("PASS " if cond else "FAIL ") + name, extra, flush=True)

i = np.arange(10000)
t = TDigest(100, mode="reference"); t.update(np.arange(1000.0)); ok("kat_a", sha(t) == G["kat_a"]["sha"], len(t.centroids()))
x = ((i * 7919) % 10007) / 10007.0; w = 1.0 + (i % 5) * 0.25
t = TDigest(100, mode="reference"); t.update(x, w); ok("kat_b", sha(t) == G["kat_b"]["sha"], len(t.centroids()))
q = t.quantile([0.001, 0.01, 0.5, 0.99, 0.999]); ok("kat_b_quantile", [float(v).hex() for v in q] == G["kat_b"]["quantile"])
c = t.cdf([0.25, 0.5, 0.75]); ok("kat_b_cdf", [float(v).hex() for v in c] == G["kat_b"]["cdf"])
x = (i * 31 % 97).astype(float)
t = TDigest(200, mode="reference"); t.update(x); ok("kat_c", sha(t) == G["kat_c"]["sha"], len(t.centroids()))
parts = []
for k in range(8):
    p = TDigest(100, mode="reference"); p.update(((i[:2000] * 7919 + k * 1237) % 10007) / 10007.0 + k); parts.append(p)
t = TDigest(100); t.merge(*parts); ok("kat_d", sha(t) == G["kat_d"]["sha"], len(t.centroids()))
t0 = time.time()
x = np.random.default_rng(0).uniform(0, 1, 10**6)
t = TDigest(100, mode="reference"); t.update(x); s = sha(t); dt = time.time() - t0
ok("config1", s == G["config1"]["sha"], "n=%d %.3fs" % (len(t.centroids()), dt))
ok("config1_q", [float(v).hex() for v in t.quantile([0.001, 0.01, 0.5, 0.99, 0.999])] == G["config1"]["quantile"])
# random vs oracle incl weights
rng = np.random.default_rng(5)
for c_ in [20, 100, 200, 1000]:
    x = rng.lognormal(size=30000); w = rng.uniform(0.5, 1.5, 30000)
    a = TDigest(c_, mode="reference"); a.update(x, w)
    b = oracle.restated.TDigest(c_); b.update(x, w)
    ok("weighted_c%d" % c_, a.centroids().tobytes() == b.centroids().tobytes() and a.size() == b.size() and a.min() == b.min())
    qs = rng.uniform(-0.1, 1.1, 1000)
    ok("quantile_c%d" % c_, a.quantile(qs).tobytes() == b.quantile(qs).tobytes())
    xs = np.concatenate([rng.lognormal(size=1000), b.centroids()["mean"]])
    ok("cdf_c%d" % c_, a.cdf(xs).tobytes() == b.cdf(xs).tobytes())
# grouped
vals = rng.normal(size=(2000, 1000))
g = TDigestGroup(2000, 100); g.update(vals)
bad = 0
for r in range(0, 2000, 97):
    b = oracle.restated.TDigest(100); b.update(vals[r])
    bad += g.centroids(r).tobytes() != b.centroids().tobytes()
ok("grouped", bad == 0)
qq = g.quantile([0.1, 0.5, 0.9]); b = oracle.restated.TDigest(100); b.update(vals[5]); ok("grouped_q", qq[5].tobytes() == b.quantile([0.1, 0.5, 0.9]).tobytes())
# tree merge
g = TDigestGroup(37, 200); v2 = rng.normal(size=(37, 2000)); g.update(v2)
tm = g.merge_tree()
ds = []
for r in range(37):
    b = oracle.restated.TDigest(200); b.update(v2[r]); ds.append(b)
s_ = 1
while s_ < 37:
    for k in range(0, 37, 2 * s_):
        if k + s_ < 37: ds[k].merge(ds[k + s_])
    s_ *= 2
ok("tree_merge", tm.centroids().tobytes() == ds[0].centroids().tobytes(), len(tm.centroids()))
# stats
xs = rng.normal(3, 2, 5000)
a = SummaryStats(); a.update(xs); b = oracle.restated.SummaryStats(); b.update(xs)
ok("stats_seq", a.__getstate__() == b.__getstate__())
xs = rng.normal(3, 2, 10**6); a = SummaryStats(); a.update(xs)
import scipy.stats
ok("stats_par", abs(a.mean() - xs.mean()) < 1e-10 and abs(a.var() - xs.var()) < 1e-9 and abs(a.kurt() - scipy.stats.kurtosis(xs)) < 1e-8, "%g %g" % (a.kurt(), scipy.stats.kurtosis(xs)))
# spsv
items = rng.zipf(1.3, 20000) % 1000
a = SpaceSaving(50, "i8"); a.update(items); b = oracle.restated.SpaceSaving(50); b.update(items)
ok("spsv", a.counters().tobytes() == b.counters().tobytes())
a2 = SpaceSaving(50, "i8"); a2.update(items[::-1].copy()); b2 = oracle.restated.SpaceSaving(50); b2.update(items[::-1].copy())
a.merge(a2); b.merge(b2); ok("spsv_merge", a.counters().tobytes() == b.counters().tobytes())
# parallel mode
n = 10**7
x = rng.lognormal(size=n); w = rng.uniform(0.5, 1.5, n)
t0 = time.time(); t = TDigest(100, mode="parallel"); t.update(x, w); nc = len(t.centroids()); dt = time.time() - t0
order = np.argsort(x); cw = np.cumsum(w[order]); tot = cw[-1]
qs = np.array([.001, .01, .1, .3, .5, .7, .9, .99, .999])
est = t.quantile(qs)
def wrank(v):
    k = np.searchsorted(x[order], v, side="left"); k2 = np.searchsorted(x[order], v, side="right")
    below = cw[k - 1] if k > 0 else 0.0; upto = cw[k2 - 1] if k2 > 0 else 0.0
    return (below + upto) / (2 * tot)
err = max(abs(wrank(v) - q_) for v, q_ in zip(est, qs))
ref = oracle.restated.TDigest(100); ref.update(x, w); err_ref = max(abs(wrank(v) - q_) for v, q_ in zip(ref.quantile(qs), qs))
ok("parallel_host", abs(t.size() - w.sum()) < 1e-6 * w.sum() and t.min() == x.min() and t.max() == x.max(), "nc=%d err=%.2e ref_err=%.2e %.2fs" % (nc, err, err_ref, dt))
# device-resident parallel + timing
for n in [10**7, 10**8, 10**9]:
    dx = DeviceArray((n,)); dw = DeviceArray((n,))
    lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None); lib.crick_device_synchronize()
    t = TDigest(100, mode="parallel")
    for rep in range(3):
        t = TDigest(100, mode="parallel")
        lib.crick_device_synchronize(); t0 = time.time()
        t.update(dx, dw); lib.crick_device_synchronize(); dt = time.time() - t0
    print("device parallel n=%d: %.3f ms -> %.1f Gval/s, centroids=%d size=%.1f" % (n, dt * 1e3, n / dt / 1e9, len(t.centroids()), t.size()), flush=True)
    del dx, dw
print("launches", cb.launch_count())
