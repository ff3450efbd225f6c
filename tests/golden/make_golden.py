"""tests/golden/make_golden.py -- regenerates tests/golden/golden.json and golden_arrays.npz.

Runs in the BUILD container only: it imports the reference itself (a scratch
copy of /root/reference built with its own `setup.py build_ext --inplace`,
SURVEY.md Appendix B) and records input recipes + outputs as fixtures.  The
GPU box never runs this; tests read the committed fixtures.

    cp -r /root/reference /tmp/oracle2 && chmod -R u+w /tmp/oracle2
    (cd /tmp/oracle2 && python3 setup.py build_ext --inplace)
    python tests/golden/make_golden.py /tmp/oracle2
"""
import hashlib
import json
import os
import pickle
import sys

import numpy as np

ref_dir = sys.argv[1] if len(sys.argv) > 1 else "/tmp/oracle2"
sys.path.insert(0, ref_dir)
import crick  # noqa: E402  (the reference)

assert os.path.dirname(crick.__file__).startswith(ref_dir)
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(t):
    return hashlib.sha256(t.centroids().tobytes()).hexdigest()


def hexs(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


G = {}
A = {}
i = np.arange(10000)
QS = [0.001, 0.01, 0.5, 0.99, 0.999]


def td_record(name, t, cdf_x):
    c = t.centroids()
    A[name + "_centroids"] = c.view(np.float64).reshape(-1, 2)
    G[name] = dict(n=len(c), size=float(t.size()).hex(), min=float(t.min()).hex(),
                   max=float(t.max()).hex(), sha=sha(t), quantile=hexs(t.quantile(QS)),
                   cdf_x=hexs(cdf_x), cdf=hexs(t.cdf(cdf_x)))


# KAT A-D
t = crick.TDigest(100); t.update(np.arange(1000.0)); td_record("kat_a", t, [0, 10.5, 499.5, 998.9, 999])
x = ((i * 7919) % 10007) / 10007.0; w = 1.0 + (i % 5) * 0.25
t = crick.TDigest(100); t.update(x, w); td_record("kat_b", t, [0.25, 0.5, 0.75])
x = (i * 31 % 97).astype(float)
t = crick.TDigest(200); t.update(x); td_record("kat_c", t, [0, 48, 48.5, 96])
parts = []
for k in range(8):
    p = crick.TDigest(100); p.update(((i[:2000] * 7919 + k * 1237) % 10007) / 10007.0 + k); parts.append(p)
t = crick.TDigest(100); t.merge(*parts); td_record("kat_d", t, [0.5, 4, 7.5])

# config 1 (BASELINE.json configs[0]): 1e6 uniform, seeded
x = np.random.default_rng(0).uniform(0, 1, 10**6)
t = crick.TDigest(100); t.update(x); td_record("config1", t, [0.1, 0.5, 0.9])

# config 2 shape, scaled: lognormal + weights, c=100
rng = np.random.default_rng(1234)
x = rng.lognormal(0, 1, 200000); w = rng.uniform(0.5, 1.5, 200000)
t = crick.TDigest(100); t.update(x, w); td_record("config2_small", t, [0.5, 1.0, 3.0])

# K-order iteration: non-contiguous inputs
rng = np.random.default_rng(7)
base = rng.normal(size=(60, 50))
wv = rng.uniform(0.5, 2, size=(60, 50))
korder = {}
for nm, xx, ww in [("rev", base.ravel()[::-1], 1), ("fortran", np.asfortranarray(base), 1),
                   ("transposed", base.T, 1), ("strided", base.ravel()[::2], 1),
                   ("rev_fw", base.ravel()[::-1], wv.ravel()), ("bcast_w", base, wv[0]),
                   ("t_w", base.T, wv.T), ("scalar_x", 1.5, wv.ravel()[:100])]:
    t = crick.TDigest(100); t.update(xx, ww); korder[nm] = sha(t)
G["korder"] = korder
A["korder_base"] = base; A["korder_w"] = wv

# scale, incl. the compaction quirk (T9)
t = crick.TDigest(); t.update([1, 2, 3, 4, 5], [1, 1000, 1, 10000, 1])
t2 = t.scale(np.finfo("f8").eps)
A["scale_quirk_centroids"] = t2.centroids().view(np.float64).reshape(-1, 2)
t = crick.TDigest(); t.update(1, [1.0, 2.0, 3.0]); A["scalar_x_centroids"] = t.centroids().view(np.float64).reshape(-1, 2)
G["compression_clip"] = [crick.TDigest(5).compression, crick.TDigest(5000).compression, crick.TDigest(100.5).compression]

# pickles produced by the reference (format compatibility)
x = ((i * 7919) % 10007) / 10007.0
t = crick.TDigest(100); t.update(x)
G["pickle_tdigest"] = pickle.dumps(t, protocol=2).hex()
s = crick.SummaryStats(); s.update(x)
G["pickle_stats"] = pickle.dumps(s, protocol=2).hex()
sp = crick.SpaceSaving(5, "i8"); sp.update([1, 1, 2, 3, 3, 3])
G["pickle_spsv_i8"] = pickle.dumps(sp, protocol=2).hex()
sp = crick.SpaceSaving(5, "f8"); sp.update([1.5, 1.5, 2.0])
G["pickle_spsv_f8"] = pickle.dumps(sp, protocol=2).hex()

# SummaryStats KATs S, S', quirks
def st(s):
    g = s.__getstate__()
    return dict(state=[int(g[0])] + hexs(g[1:7]) + [bool(g[7])] + hexs([g[8]]),
                mean=float(s.mean()).hex(), var=float(s.var()).hex(), std=float(s.std()).hex(),
                skew=float(s.skew()).hex(), kurt=float(s.kurt()).hex(),
                skew_unbiased=float(s.skew(bias=False)).hex(),
                kurt_pearson_unbiased=float(s.kurt(fisher=False, bias=False)).hex())
s = crick.SummaryStats(); s.update(((i * 7919) % 10007) / 10007.0); G["stats_s"] = st(s)
s = crick.SummaryStats(); s.update((i * 31 % 97).astype(float)); G["stats_s2"] = st(s)
s = crick.SummaryStats(); s.update([-5.0, -3.0]); G["stats_negmax"] = st(s)
s = crick.SummaryStats(); s.add(10, 2); G["stats_add_count"] = st(s)
s = crick.SummaryStats(); s.update([1.0, np.nan, 1.0]); G["stats_nan_mid"] = st(s)
s = crick.SummaryStats(); s.update([np.nan, 1.0, 1.0]); G["stats_nan_first"] = st(s)
rng = np.random.default_rng(5)
xs = rng.normal(3, 2, 5000); cs = rng.integers(1, 6, 5000)
A["stats_w_x"] = xs; A["stats_w_c"] = cs
s = crick.SummaryStats(); s.update(xs, cs); G["stats_weighted"] = st(s)
a = crick.SummaryStats(); a.update(xs[:2000]); b = crick.SummaryStats(); b.update(xs[2000:] + 1); a.merge(b)
G["stats_merge"] = st(a)

# SpaceSaving
def ctr(sp):
    c = sp.counters()
    return [[(float(v).hex() if sp.dtype.kind == "f" else int(v)) for v in c["item"]],
            [int(v) for v in c["count"]], [int(v) for v in c["error"]]]
sp = crick.SpaceSaving(2, "i8"); sp.add(1, 5); sp.add(2, 3); sp.add(3, 100); G["spsv_evict_ignores_w"] = ctr(sp)
sp = crick.SpaceSaving(5, "f8"); sp.update([0.0, -0.0, 0.0, np.nan, np.nan]); G["spsv_f8_bits"] = ctr(sp)
steps = []
sp = crick.SpaceSaving(3, "i8")
for it, c in [(1, 1), (2, 1), (1, 1), (3, 2), (4, 1), (4, 1), (5, 3), (2, 2), (1, 1)]:
    sp.add(it, c); steps.append(ctr(sp))
G["spsv_steps"] = steps
rs = np.random.RandomState(42)
data = (rs.gamma(0.1, 0.1, 10000) * 1000).astype("i8")  # few hundred distinct ints, skewed
A["spsv_data"] = data
sp = crick.SpaceSaving(20, "i8"); sp.update(data); G["spsv_cap20"] = ctr(sp)
sp_big = crick.SpaceSaving(2000, "i8"); sp_big.update(data); G["spsv_exact"] = ctr(sp_big)
a = crick.SpaceSaving(20, "i8"); a.update(data[:5000]); b = crick.SpaceSaving(20, "i8"); b.update(data[5000:])
a.merge(b); G["spsv_merge"] = ctr(a)
cw = rs.randint(1, 4, 10000)
A["spsv_counts"] = cw
sp = crick.SpaceSaving(50, "i8"); sp.update(data, cw); G["spsv_weighted"] = ctr(sp)

G["_meta"] = dict(numpy=np.__version__, reference_version=crick.__version__,
                  note="generated by tests/golden/make_golden.py from the built reference")
with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(G, f, indent=1, sort_keys=True)
np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **A)
print("wrote", len(G), "records,", len(A), "arrays")
