"""CPU: pins the oracle (oracle/crick_oracle.c) against the reference's golden vectors
(tests/golden, generated from the built reference) and, when oracle/_ref was built, against the
unmodified reference C sources bit for bit."""
import numpy as np
import pytest

import oracle
from helpers import (QS5, check_digest_against_golden, hexs, kat_d_parts, kat_inputs, sha)

R = oracle.restated
needs_ref = pytest.mark.skipif(oracle.reference is None, reason="oracle/_ref not built")


@pytest.mark.parametrize("name", ["kat_a", "kat_b", "kat_c"])
def test_tdigest_kats(golden, name):
    g, _ = golden
    c, x, w, cdf_x = kat_inputs()[name]
    t = R.TDigest(c)
    t.update(x, w)
    check_digest_against_golden(t, g[name], cdf_x)


def test_tdigest_kat_d_merge(golden):
    g, _ = golden
    parts = []
    for x in kat_d_parts():
        p = R.TDigest(100)
        p.update(x)
        parts.append(p)
    t = R.TDigest(100)
    t.merge(*parts)
    check_digest_against_golden(t, g["kat_d"], [0.5, 4, 7.5])


def test_config1_and_config2_small(golden):
    g, _ = golden
    t = R.TDigest(100)
    t.update(np.random.default_rng(0).uniform(0, 1, 10**6))
    check_digest_against_golden(t, g["config1"], [0.1, 0.5, 0.9])
    rng = np.random.default_rng(1234)
    x = rng.lognormal(0, 1, 200000)
    w = rng.uniform(0.5, 1.5, 200000)
    t = R.TDigest(100)
    t.update(x, w)
    check_digest_against_golden(t, g["config2_small"], [0.5, 1.0, 3.0])


def test_scale_quirk_and_scalar_x(golden):
    g, a = golden
    t = R.TDigest()
    t.update([1, 2, 3, 4, 5], [1, 1000, 1, 10000, 1])
    t2 = t.scale(np.finfo("f8").eps)
    assert t2.centroids().view(np.float64).reshape(-1, 2).tobytes() == a["scale_quirk_centroids"].tobytes()
    assert g["compression_clip"] == [R.TDigest(5).compression, R.TDigest(5000).compression,
                                     R.TDigest(100.5).compression]


def test_empty_and_single():
    t = R.TDigest()
    assert t.size() == 0 and len(t.centroids()) == 0
    assert np.isnan(t.min()) and np.isnan(t.quantile([0.5])[0]) and np.isnan(t.cdf([0.5])[0])
    t.add(10)
    assert t.quantile([0, 0.5, 1]).tolist() == [10, 10, 10]
    assert t.cdf([9, 10, 11]).tolist() == [0, 0.5, 1]


def _st(s):
    g = s.__getstate__()
    return [int(g[0])] + hexs(g[1:7]) + [bool(g[7])] + hexs([g[8]])


def test_stats_goldens(golden):
    g, a = golden
    i = np.arange(10000)
    cases = {
        "stats_s": (((i * 7919) % 10007) / 10007.0, 1),
        "stats_s2": ((i * 31 % 97).astype(float), 1),
        "stats_negmax": (np.array([-5.0, -3.0]), 1),
        "stats_nan_mid": (np.array([1.0, np.nan, 1.0]), 1),
        "stats_nan_first": (np.array([np.nan, 1.0, 1.0]), 1),
        "stats_weighted": (a["stats_w_x"], a["stats_w_c"]),
    }
    for name, (x, c) in cases.items():
        s = R.SummaryStats()
        s.update(x, c)
        rec = g[name]
        assert _st(s) == rec["state"], name
        for k, v in [("mean", s.mean()), ("var", s.var()), ("std", s.std()), ("skew", s.skew()),
                     ("kurt", s.kurt()), ("skew_unbiased", s.skew(bias=False)),
                     ("kurt_pearson_unbiased", s.kurt(fisher=False, bias=False))]:
            assert float(v).hex() == rec[k] or (np.isnan(v) and rec[k] == "nan"), (name, k)
    s = R.SummaryStats()
    s.add(10, 2)
    assert _st(s) == g["stats_add_count"]["state"]
    x = a["stats_w_x"]
    p = R.SummaryStats(); p.update(x[:2000])
    q = R.SummaryStats(); q.update(x[2000:] + 1)
    p.merge(q)
    assert _st(p) == g["stats_merge"]["state"]


def _ctr(sp):
    c = sp.counters()
    return [[int(v) for v in c["item"]], [int(v) for v in c["count"]], [int(v) for v in c["error"]]]


def test_spsv_goldens(golden):
    g, a = golden
    sp = R.SpaceSaving(2); sp.add(1, 5); sp.add(2, 3); sp.add(3, 100)
    assert _ctr(sp) == g["spsv_evict_ignores_w"]
    sp = R.SpaceSaving(3)
    for (it, c), want in zip([(1, 1), (2, 1), (1, 1), (3, 2), (4, 1), (4, 1), (5, 3), (2, 2), (1, 1)],
                             g["spsv_steps"]):
        sp.add(it, c)
        assert _ctr(sp) == want
    data = a["spsv_data"]
    sp = R.SpaceSaving(20); sp.update(data); assert _ctr(sp) == g["spsv_cap20"]
    sp = R.SpaceSaving(2000); sp.update(data); assert _ctr(sp) == g["spsv_exact"]
    p = R.SpaceSaving(20); p.update(data[:5000]); q = R.SpaceSaving(20); q.update(data[5000:])
    p.merge(q); assert _ctr(p) == g["spsv_merge"]
    sp = R.SpaceSaving(50); sp.update(data, a["spsv_counts"]); assert _ctr(sp) == g["spsv_weighted"]
    # float items are tracked by bit pattern: 0.0 / -0.0 distinct, equal-bit NaNs coincide
    bits = np.array([0.0, -0.0, 0.0, np.nan, np.nan]).view("i8")
    sp = R.SpaceSaving(5); sp.update(bits)
    got = sp.counters()
    assert [float(v).hex() for v in got["item"].view("f8")] == g["spsv_f8_bits"][0]
    assert [int(v) for v in got["count"]] == g["spsv_f8_bits"][1]


@pytest.mark.gpu          # (no GPU needed: runs where oracle/_ref travels with the snapshot, i.e. the GPU tier)
@needs_ref
def test_restated_vs_reference_randomised():
    F = oracle.reference
    rng = np.random.default_rng(42)
    for c in [20, 100, 200, 1000, 100.5]:
        for x, w in [(rng.uniform(0, 1, 20000), 1.0),
                     (rng.lognormal(size=20000), rng.uniform(0.5, 1.5, 20000)),
                     (rng.integers(0, 50, 20000).astype(float), 1.0),
                     (np.where(rng.integers(0, 2, 5000) == 0, 0.0, -0.0), rng.uniform(0.1, 3, 5000)),
                     (np.sort(rng.normal(size=20000)), 1.0),
                     (np.where(rng.uniform(size=5000) < 0.1, np.nan, rng.normal(size=5000)), 1.0)]:
            a, b = R.TDigest(c), F.TDigest(c)
            a.update(x, w); b.update(x, w)
            assert a.raw() == b.raw()
            assert a.centroids().tobytes() == b.centroids().tobytes()
            q = np.concatenate([rng.uniform(-0.1, 1.1, 300), [0, 1, 1e-12, 1 - 1e-12]])
            assert a.quantile(q).tobytes() == b.quantile(q).tobytes()
            cen = b.centroids()["mean"]
            xs = np.concatenate([rng.uniform(np.nanmin(x) - 1, np.nanmax(x) + 1, 300), cen,
                                 np.nextafter(cen, np.inf), [b.min(), b.max()]])
            assert a.cdf(xs).tobytes() == b.cdf(xs).tobytes()
    pa, pb = [], []
    for k in range(30):
        x = rng.normal(k * 0.1, 1, 3000)
        a, b = R.TDigest(200), F.TDigest(200)
        a.update(x); b.update(x); pa.append(a); pb.append(b)
    A, B = R.TDigest(200), F.TDigest(200)
    A.merge(*pa); B.merge(*pb)
    assert A.centroids().tobytes() == B.centroids().tobytes() and A.raw() == B.raw()
    assert A.scale(0.5).centroids().tobytes() == B.scale(0.5).centroids().tobytes()


@pytest.mark.gpu
@needs_ref
def test_restated_stats_spsv_vs_reference():
    F = oracle.reference
    rng = np.random.default_rng(43)
    for x, c in [(rng.normal(size=10000), 1), (rng.normal(size=3000), rng.integers(1, 5, 3000)),
                 (np.where(rng.uniform(size=1000) < 0.1, np.nan, 1.0), 1),
                 (rng.normal(size=2_500_000), 1)]:  # last: int64 products overflow, identically
        a, b = R.SummaryStats(), F.SummaryStats()
        a.update(x, c); b.update(x, c)
        assert _st(a) == _st(b)
    for trial in range(20):
        cap = int(rng.integers(2, 40)); n = int(rng.integers(1, 3000))
        items = rng.zipf(1.3, n) % 100
        cnt = rng.integers(1, 5, n) if trial % 2 else 1
        a, b = R.SpaceSaving(cap), F.SpaceSaving(cap)
        a.update(items, cnt); b.update(items, cnt)
        assert a.counters().tobytes() == b.counters().tobytes()
        a2, b2 = R.SpaceSaving(cap), F.SpaceSaving(cap)
        items2 = rng.zipf(1.3, n) % 120
        a2.update(items2); b2.update(items2)
        a.merge(a2); b.merge(b2)
        assert a.counters().tobytes() == b.counters().tobytes()
