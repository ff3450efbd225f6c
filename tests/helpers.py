"""Shared test helpers (test infrastructure: may import oracle/)."""
import hashlib
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
QS9 = np.array([0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.99, 0.999])
QS5 = [0.001, 0.01, 0.5, 0.99, 0.999]


def load_golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        g = json.load(f)
    a = np.load(os.path.join(HERE, "golden", "golden_arrays.npz"))
    return g, a


def sha(t):
    return hashlib.sha256(t.centroids().tobytes()).hexdigest()


def hexs(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def kat_inputs():
    """Input recipes of tests/golden/make_golden.py."""
    i = np.arange(10000)
    return {
        "kat_a": (100, np.arange(1000.0), 1.0, [0, 10.5, 499.5, 998.9, 999]),
        "kat_b": (100, ((i * 7919) % 10007) / 10007.0, 1.0 + (i % 5) * 0.25, [0.25, 0.5, 0.75]),
        "kat_c": (200, (i * 31 % 97).astype(float), 1.0, [0, 48, 48.5, 96]),
    }


def kat_d_parts():
    i = np.arange(10000)
    return [((i[:2000] * 7919 + k * 1237) % 10007) / 10007.0 + k for k in range(8)]


def check_digest_against_golden(t, rec, cdf_x):
    c = t.centroids()
    assert len(c) == rec["n"]
    assert sha(t) == rec["sha"]
    assert float(t.size()).hex() == rec["size"]
    assert float(t.min()).hex() == rec["min"]
    assert float(t.max()).hex() == rec["max"]
    assert hexs(t.quantile(QS5)) == rec["quantile"]
    assert hexs(t.cdf(cdf_x)) == rec["cdf"]


def weighted_rank_fn(x, w):
    """Returns rank(v) = (weight below v + weight up to v) / (2 total), the weighted analogue of
    the reference tests' quantiles_to_q (crick/tests/test_tdigest.py:52-56)."""
    order = np.argsort(x, kind="stable")
    xs = x[order]
    cw = np.concatenate([[0.0], np.cumsum(np.broadcast_to(w, x.shape)[order])])
    tot = cw[-1]

    def rank(v):
        v = np.atleast_1d(v)
        lo = np.searchsorted(xs, v, side="left")
        hi = np.searchsorted(xs, v, side="right")
        return (cw[lo] + cw[hi]) / (2 * tot)
    return rank


def quantiles_to_q(data, quant):
    n = len(data)
    s = np.sort(data)
    lo = np.searchsorted(s, quant, side="left")
    hi = np.searchsorted(s, quant, side="right")
    return (lo + hi) / (2 * n)


def q_to_x(data, q):
    n = len(data)
    s = np.sort(data)
    x = np.empty_like(q)
    for i in range(len(q)):
        ix = n * q[i] - 0.5
        index = int(ix)
        p = ix - index
        x[i] = s[index] * (1 - p) + s[index + 1] * p
    return x


def kscale(c, q):
    q = min(q, 1.0)
    return c * (math.asin(2 * q - 1) + math.pi / 2) / math.pi


def restate_parallel_items(edges, bw, bs, mn, mx, comp):
    """CPU restatement of k_finalize's bucket -> weighted-point step
    (crick_b200/csrc/tdigest_parallel.cu), for the parallel-mode parity tests."""
    E = len(edges)
    B = E + 1
    assert len(bw) == B
    Wb = 0.0
    for v in bw:
        Wb += v
    items = []
    before = 0.0
    for b in range(B):
        w_ = bw[b]
        if w_ > 0.0:
            single = 1 <= b <= E - 1 and edges[b] == np.nextafter(edges[b - 1], np.inf)
            p = 1
            if single:
                span = kscale(comp, (before + w_) / Wb) - kscale(comp, before / Wb)
                p = max(1, int(math.ceil(span * 2.0)))
            lo = mn if b == 0 else edges[b - 1]
            hi = mx if b == E else np.nextafter(edges[b], -np.inf)
            mean = bs[b] / w_
            if p > 1 or single:
                mean = edges[b - 1]
            mean = min(max(mean, lo), hi)
            each = w_ / p
            for k in range(p):
                items.append((mean, (w_ - each * (p - 1)) if k == p - 1 else each))
        before += w_
    return np.array(items, dtype=np.float64).reshape(-1, 2)


def accepted(x, w):
    x = np.asarray(x, dtype=np.float64)
    w = np.broadcast_to(np.asarray(w, dtype=np.float64), x.shape)
    m = np.isfinite(x) & ~(w <= np.finfo("f8").eps)
    return x[m], w[m]
