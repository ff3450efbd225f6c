"""Tail accuracy of the parallel k-bucket mode against the reference's own digest on the same data:
weighted rank error at q in {1e-5, 1e-4, 1e-3, 1 - 1e-3, 1 - 1e-4, 1 - 1e-5}.
  python tests/tools/tail_probe.py [n] [kind: 0 lognormal+weights, 5 gamma(0.1)] [compression]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib
import oracle

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
comp = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
R = oracle.reference if oracle.reference is not None else oracle.restated
QT = np.array([1e-5, 1e-4, 1e-3, 1 - 1e-3, 1 - 1e-4, 1 - 1e-5])

if kind == 0:
    dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
    lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None)
    x, w = dx.to_numpy(), dw.to_numpy()
else:
    rng = np.random.default_rng(7)
    x = rng.gamma(0.1, 1.0, n)
    w = np.ones(n)
    dx, dw = cb.DeviceArray.from_numpy(x), None

def rank_of(v):
    below = w[x < v].sum()
    upto = w[x <= v].sum()
    return (below + upto) / (2.0 * tot)
tot = w.sum()
out = {"n": n, "kind": kind, "compression": comp}
for seed_shift in range(1):
    t = cb.TDigest(comp, mode="parallel")
    t.update(dx, dw) if dw is not None else t.update(dx)
    est = t.quantile(QT)
    out["ours"] = [abs(rank_of(v) - q) for v, q in zip(est, QT)]
    out["ours_centroids"] = len(t.centroids())
    c = t.centroids()
    out["ours_first_last_weights"] = [c["weight"][:3].tolist(), c["weight"][-3:].tolist()]
t0 = time.perf_counter()
r = R.TDigest(comp)
r.update(x, w)
out["ref_seconds"] = time.perf_counter() - t0
est = r.quantile(QT)
out["ref"] = [abs(rank_of(v) - q) for v, q in zip(est, QT)]
c = r.centroids()
out["ref_centroids"] = len(c)
out["ref_first_last_weights"] = [c["weight"][:3].tolist(), c["weight"][-3:].tolist()]
out["q"] = QT.tolist()
print(json.dumps(out), flush=True)
