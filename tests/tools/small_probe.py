"""Rank error and time of the parallel path on small inputs, against the oracle (test tooling)."""
import sys, numpy as np, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import crick_b200 as cb, oracle
from crick_b200 import _lib
from helpers import QS9, weighted_rank_fn
_lib.PARALLEL_MIN = 1024
rng = np.random.default_rng(5)
for dist in ("lognormal", "normal", "gamma"):
    for n in (4096, 8192, 20000, 50000):
        x = {"lognormal": rng.lognormal(size=n), "normal": rng.normal(size=n), "gamma": rng.gamma(0.1, 0.1, n)}[dist]
        rank = weighted_rank_fn(x, 1.0)
        errs = []
        for rep in range(3):
            t = cb.TDigest(100, mode="parallel")
            # bypass the wrapper's fallback by calling the library directly
            from crick_b200._lib import lib, check
            check(lib.crick_tdigest_update(t._h, x.ctypes.data, None, 1.0, n, 0, 0, 2, None))
            errs.append(np.abs(rank(t.quantile(QS9)) - QS9).max())
        r = oracle.restated.TDigest(100); r.update(x)
        er = np.abs(rank(r.quantile(QS9)) - QS9).max()
        t0 = time.perf_counter(); t = cb.TDigest(100, mode="reference"); t.update(x); t.centroids(); tr = time.perf_counter() - t0
        t0 = time.perf_counter(); t = cb.TDigest(100); check(lib.crick_tdigest_update(t._h, x.ctypes.data, None, 1.0, n, 0, 0, 2, None)); t.centroids(); tp = time.perf_counter() - t0
        print("%-9s n=%6d  parallel err %.2e  ref err %.2e   gpu-ref %.1f ms  gpu-par %.2f ms  nc=%d" % (dist, n, max(errs), er, tr * 1e3, tp * 1e3, len(t.centroids())))
