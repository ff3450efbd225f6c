import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import crick_b200 as cb, oracle
from helpers import QS9, weighted_rank_fn
rng = np.random.default_rng(5)
worse = 0
for dist in ("lognormal", "normal", "gamma", "uniform", "ties"):
    for n in (2048, 2500, 3000, 4096, 6000):
        pe, re_ = [], []
        for rep in range(8):
            x = {"lognormal": rng.lognormal(size=n), "normal": rng.normal(size=n), "gamma": rng.gamma(0.1, 0.1, n),
                 "uniform": rng.uniform(size=n), "ties": rng.integers(0, 50, n).astype(float)}[dist]
            rank = weighted_rank_fn(x, 1.0)
            t = cb.TDigest(100, mode="parallel"); t.update(x)
            r = oracle.restated.TDigest(100); r.update(x)
            if dist == "ties":
                # rank error is ill-defined on ties: compare the quantiles themselves with numpy's
                qp, qr = t.quantile(QS9), r.quantile(QS9)
                qt = np.quantile(x, QS9)
                pe.append(np.abs(qp - qt).max()); re_.append(np.abs(qr - qt).max())
            else:
                pe.append(np.abs(rank(t.quantile(QS9)) - QS9).max()); re_.append(np.abs(rank(r.quantile(QS9)) - QS9).max())
        flag = "" if np.mean(pe) <= 1.15 * np.mean(re_) + 1e-12 else "  <-- worse"
        worse += bool(flag)
        print("%-9s n=%5d  parallel %.2e (max %.2e)   reference %.2e (max %.2e)%s" % (dist, n, np.mean(pe), np.max(pe), np.mean(re_), np.max(re_), flag), flush=True)
print("worse cases:", worse)
