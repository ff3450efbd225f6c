"""GPU parity tests for SummaryStats (stats_stubs.c) and SpaceSaving (space_saving_stubs.c.in),
through the C ABI, against the oracle and the reference's own known-answer tests
(crick/tests/test_stats.py, crick/tests/test_space_saving.py).

Bar: SummaryStats bit-exact on the sequential replay (auto below 256 values per call, forced
here for the 10000-value known answers), moments within 1e-12 relative / `atol 1e-8`-style
tolerance vs NumPy/SciPy on the parallel path (the reference's own test tolerance,
test_stats.py:20-21); SpaceSaving counters bit-exact, including list order.
"""
import pickle

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
R = oracle.restated


@pytest.fixture(scope="module")
def cb():
    import crick_b200
    if crick_b200.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return crick_b200


# ---- SummaryStats -------------------------------------------------------------------------
@pytest.fixture
def stats_replay():
    """Force the sequential replay of stats_do_update (bit-exact with the reference); auto mode
    uses it below 256 values per call only, because it costs ~1.5 us per element."""
    from crick_b200._lib import lib
    prev = lib.crick_stats_kernel(0)
    yield
    lib.crick_stats_kernel(prev)


def test_stats_golden_state(cb, golden, stats_replay):
    g, _ = golden
    i = np.arange(10000)
    s = cb.SummaryStats()
    s.update(((i * 7919) % 10007) / 10007.0)
    o = R.SummaryStats()
    o.update(((i * 7919) % 10007) / 10007.0)
    assert s.__getstate__() == o.__getstate__()
    assert (s.var(), s.skew(), s.kurt()) == (o.var(), o.skew(), o.kurt())
    # SURVEY Appendix A, KAT S
    assert s.__getstate__()[0] == 10000 and s.sum() == 5000.157689617268
    assert s.var() == 0.08334254276727161 and s.kurt() == -1.20010745225976


@pytest.mark.parametrize("case", ["normal_nan", "empty", "one", "duplicate", "different", "ints"])
def test_stats_sequential_bit_exact(cb, case, stats_replay):
    rng = np.random.default_rng(0)
    x = {"normal_nan": np.where(rng.uniform(size=10000) < 0.1, np.nan, rng.normal(50, 10, 10000)),
         "empty": np.array([]), "one": np.array([1.5]), "duplicate": np.array([1.5] * 10),
         "different": np.array([1.0, 2.0]), "ints": rng.integers(-50, 50, 3000)}[case]
    s, o = cb.SummaryStats(), R.SummaryStats()
    s.update(x)
    o.update(x)
    a, b = s.__getstate__(), o.__getstate__()
    assert np.array_equal(np.array(a, dtype=float), np.array(b, dtype=float), equal_nan=True)
    for f in ("mean", "var", "std", "skew", "kurt"):
        assert np.array_equal(getattr(s, f)(), getattr(o, f)(), equal_nan=True), f
    assert np.array_equal(s.var(ddof=1), o.var(ddof=1), equal_nan=True)
    assert np.array_equal(s.skew(bias=False), o.skew(False), equal_nan=True)
    assert np.array_equal(s.kurt(fisher=False, bias=False), o.kurt(False, False), equal_nan=True)


@pytest.mark.parametrize("n", [256, 1024, 3000, 10000, 60000])
def test_stats_parallel_pass_on_small_arrays(cb, n):
    """Auto mode from 256 values: the parallel pass must agree with the reference to ~1e-12
    relative in every moment and exactly in count / min / max / sum of integers, NaNs skipped
    and the homogeneous / first-value bookkeeping intact."""
    rng = np.random.default_rng(n)
    x = rng.normal(50, 10, n)
    x[rng.integers(0, n, 7)] = np.nan
    for data in (x, rng.integers(-50, 50, n), np.full(n, 2.5)):
        s, o = cb.SummaryStats(), R.SummaryStats()
        s.update(data)
        o.update(data)
        a, b = s.__getstate__(), o.__getstate__()
        assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3] and a[7] == b[7] and a[8] == b[8]
        np.testing.assert_allclose(a[1], b[1], rtol=1e-13)                       # sum
        for f in ("mean", "var", "std", "skew", "kurt"):
            np.testing.assert_allclose(getattr(s, f)(), getattr(o, f)(), rtol=1e-10, atol=1e-12, equal_nan=True, err_msg=f)
        m = o.__getstate__()
        np.testing.assert_allclose(a[4:7], m[4:7], rtol=1e-11, atol=1e-9 * abs(m[4]) if m[4] else 1e-300)


def test_stats_quirks(cb):
    s = cb.SummaryStats()
    s.update([-5.0, -3.0])
    assert s.max() == -2.2250738585072014e-308          # S0: max starts at -DBL_MIN
    s = cb.SummaryStats()
    s.add(10, 2)                                        # test_stats.py:112-116
    assert (s.count(), s.sum(), s.mean(), s.var()) == (2, 10.0, 5.0, 0.0)
    s = cb.SummaryStats()
    s.update([1.0, np.nan, 1.0])                        # S2: NaN after the first value
    assert s.__getstate__() == (2, 2.0, 1.0, 1.0, 0.0, 0.0, 0.0, False, 1.0)
    assert s.skew() == 0.0 and s.kurt() == -3.0
    s = cb.SummaryStats()
    s.update([np.nan, 1.0, 1.0])
    assert np.isnan(s.skew()) and np.isnan(s.kurt())
    assert repr(cb.SummaryStats()).startswith("SummaryStats<count=0")
    with pytest.raises(ValueError):
        s.add(1, 0)
    with pytest.raises(TypeError):
        s.update(np.array(["a"]))
    with pytest.raises(TypeError):
        s.merge(1)


def test_stats_weights_merge_pickle(cb, stats_replay):
    rng = np.random.default_rng(1)
    x = rng.normal(size=4000)
    c = rng.integers(1, 5, 4000)
    s, o = cb.SummaryStats(), R.SummaryStats()
    s.update(x, c)
    o.update(x, c)
    assert s.__getstate__() == o.__getstate__()
    parts, oparts = [], []
    for k in range(4):
        p, q = cb.SummaryStats(), R.SummaryStats()
        p.update(x[k * 1000:(k + 1) * 1000])
        q.update(x[k * 1000:(k + 1) * 1000])
        parts.append(p)
        oparts.append(q)
    m, mo = cb.SummaryStats(), R.SummaryStats()
    m.merge(*parts)
    mo.merge(*oparts)
    assert m.__getstate__() == mo.__getstate__()
    m.merge(cb.SummaryStats())                          # merge with empty is the identity
    assert m.__getstate__() == mo.__getstate__()
    m2 = pickle.loads(pickle.dumps(m))
    assert m2.__getstate__() == m.__getstate__()
    rec = m.export_record()
    z = cb.SummaryStats()
    z.merge_record(rec)
    assert z.__getstate__() == m.__getstate__()


def test_stats_parallel_path_vs_numpy_scipy(cb):
    import scipy.stats
    rng = np.random.default_rng(2)
    x = rng.normal(3, 2, 2 * 10**6)
    x[::1000] = np.nan
    for src in (x, None):
        s = cb.SummaryStats()
        s.update(cb.DeviceArray.from_numpy(x) if src is None else x)
        ok = x[~np.isnan(x)]
        assert s.count() == ok.size and s.min() == ok.min() and s.max() == ok.max()
        assert abs(s.sum() - ok.sum()) <= 1e-9 * abs(ok.sum())
        np.testing.assert_allclose(s.mean(), ok.mean(), rtol=0, atol=1e-8)
        np.testing.assert_allclose(s.var(), ok.var(), rtol=0, atol=1e-8)
        np.testing.assert_allclose(s.skew(), scipy.stats.skew(ok), rtol=0, atol=1e-8)
        np.testing.assert_allclose(s.kurt(), scipy.stats.kurtosis(ok), rtol=0, atol=1e-8)
    # where the reference's int64 products wrap (n > 2^21) it is the reference
    # that is wrong: the oracle with fp64 products agrees with us and with SciPy
    o = R.SummaryStats()
    oracle.stats_fp_products(o, True)
    o.update(x)
    np.testing.assert_allclose(s.kurt(), o.kurt(), rtol=0, atol=1e-8)
    # int64 input converts like the NumPy safe cast (2^53 + 1 rounds)
    xi = rng.zipf(1.1, 500000).astype(np.int64)
    s = cb.SummaryStats()
    s.update(cb.DeviceArray.from_numpy(xi))
    xf = xi.astype(np.float64)
    assert s.count() == xi.size and s.min() == xf.min() and s.max() == xf.max()
    assert abs(s.mean() - xf.mean()) <= 1e-9 * abs(xf.mean())


# ---- SpaceSaving ----------------------------------------------------------------------------
def test_spsv_step_by_step_kat(cb):
    """crick/tests/test_space_saving.py:67-107 and the weights KAT :164-184."""
    s = cb.SpaceSaving(capacity=3, dtype="i8")
    o = R.SpaceSaving(3)
    for it, c in [(1, 1), (2, 1), (3, 1), (2, 1), (4, 1), (4, 1), (5, 3), (1, 2), (6, 1), (2, 5)]:
        s.add(it, c)
        o.add(it, c)
        assert s.counters().tobytes() == o.counters().tobytes()
    s = cb.SpaceSaving(2, "i8")
    s.add(1, 5)
    s.add(2, 3)
    s.add(3, 100)                                       # eviction ignores the weight (:229-231)
    assert s.counters().tolist() == [(1, 5, 0), (3, 4, 3)]
    f = cb.SpaceSaving(5, "f8")
    f.update([0.0, -0.0, 0.0, np.nan, np.nan])          # keyed by bit pattern
    c = f.counters()
    assert c["count"].tolist() == [2, 2, 1] and c["error"].tolist() == [0, 0, 0]
    assert np.signbit(c["item"][2]) and np.isnan(c["item"][1])


@pytest.mark.parametrize("dtype", ["i8", "f8"])
def test_spsv_stream_vs_oracle_and_topk(cb, dtype):
    data = np.random.RandomState(42).gamma(0.1, 0.1, size=10000).round(2) * 100   # test_space_saving.py:10-11
    if dtype == "i8":
        data = data.astype("i8")
    s, o = cb.SpaceSaving(20, dtype), R.SpaceSaving(20)
    s.update(data)
    o.update(data.view("i8") if dtype == "f8" else data)
    assert s.counters().view("i8").tobytes() == o.counters().view("i8").tobytes()
    assert s.size() == s.capacity
    vals, counts = np.unique(data, return_counts=True)
    inds = counts.argsort()[::-1]                       # test_space_saving.py:16-19,34-41
    top = s.topk(10)
    assert (top["item"] == vals[inds[:10]]).all()
    assert (top["count"] == counts[inds[:10]]).all()
    true = dict(zip(vals.tolist(), counts.tolist()))
    for it, c, e in s.counters().tolist():
        assert c - e <= true.get(it, 0) <= c            # the Space-Saving guarantee
    assert (np.diff(s.counters()["count"]) <= 0).all()
    t = s.topk(3, astuples=True)
    assert len(t) == 3 and t[0].count >= t[1].count
    s2 = pickle.loads(pickle.dumps(s))
    assert s2.counters().tobytes() == s.counters().tobytes()
    a = cb.SpaceSaving(20, dtype)
    for v in data[:500]:
        a.add(v)
    b = cb.SpaceSaving(20, dtype)
    b.update(data[:500])
    assert a.counters().tobytes() == b.counters().tobytes()      # add == update


def test_spsv_exact_when_distinct_fits(cb):
    """north_star: counts bit-exact whenever #distinct <= capacity; order = (count desc,
    last-update position asc) as the reference's list gives it."""
    rng = np.random.default_rng(3)
    items = rng.zipf(1.3, 50000) % 700
    w = rng.integers(1, 4, items.size)
    s, o = cb.SpaceSaving(1000, "i8"), R.SpaceSaving(1000)
    s.update(items, w)
    o.update(items, w)
    assert s.counters().tobytes() == o.counters().tobytes()
    vals = np.unique(items)
    got = {it: c for it, c, e in s.counters().tolist()}
    assert all(got[v] == w[items == v].sum() for v in vals[:50])
    assert (s.counters()["error"] == 0).all()
    d = cb.SpaceSaving(1000, "i8")
    d.update(cb.DeviceArray.from_numpy(items.astype(np.int64)), cb.DeviceArray.from_numpy(w.astype(np.int64)))
    assert d.counters().tobytes() == s.counters().tobytes()


def test_spsv_merge_vs_oracle(cb):
    rng = np.random.default_rng(4)
    a_items = rng.zipf(1.2, 20000) % 3000
    b_items = (rng.zipf(1.2, 20000) % 3000)[::-1].copy()
    for cap_a, cap_b in [(50, 50), (50, 20), (20, 50), (5000, 5000)]:
        a, oa = cb.SpaceSaving(cap_a, "i8"), R.SpaceSaving(cap_a)
        b, ob = cb.SpaceSaving(cap_b, "i8"), R.SpaceSaving(cap_b)
        a.update(a_items)
        oa.update(a_items)
        b.update(b_items)
        ob.update(b_items)
        before = b.counters().tobytes()
        a.merge(b)
        oa.merge(ob)
        assert a.counters().tobytes() == oa.counters().tobytes(), (cap_a, cap_b)
        assert b.counters().tobytes() == before         # merge does not change its argument
    with pytest.raises(TypeError):
        a.merge(1)
    with pytest.raises(ValueError):
        a.merge(cb.SpaceSaving(5, "f8"))
    with pytest.raises(ValueError):
        cb.SpaceSaving(0)
    with pytest.raises(ValueError):
        a.add(1, 0)


def test_spsv_parallel_path(cb):
    """Streams above 262144 elements: shard summaries + the reference's merge.  Exact (counts,
    errors AND list order) whenever #distinct <= capacity; otherwise a valid Space-Saving summary:
    count - error <= true <= count, sorted, and the heavy hitters are found."""
    rng = np.random.default_rng(6)
    n = 700003
    items = (rng.zipf(1.2, n) % 900).astype(np.int64)
    w = rng.integers(1, 4, n).astype(np.int64)
    s, o = cb.SpaceSaving(1000, "i8"), R.SpaceSaving(1000)
    s.update(cb.DeviceArray.from_numpy(items), cb.DeviceArray.from_numpy(w))
    o.update(items, w)
    assert s.counters().tobytes() == o.counters().tobytes()
    s.update(items[::-1].copy())                      # a second batch on top, host path
    o.update(items[::-1].copy())
    assert s.counters().tobytes() == o.counters().tobytes()
    # more distinct items than counters
    items = rng.zipf(1.1, n).astype(np.int64)
    s = cb.SpaceSaving(1000, "i8")
    s.update(items)
    c = s.counters()
    assert len(c) == 1000 and (np.diff(c["count"]) <= 0).all()
    vals, counts = np.unique(items, return_counts=True)
    true = dict(zip(vals.tolist(), counts.tolist()))
    for it, cn, er in c.tolist():
        assert cn - er <= true.get(it, 0) <= cn
    top_true = set(vals[np.argsort(-counts, kind="stable")[:20]].tolist())
    assert top_true <= set(c["item"][:60].tolist())
    # every item whose true count exceeds the smallest counter must be present (guarantee)
    floor = c["count"].min()
    assert set(vals[counts > floor].tolist()) <= set(c["item"].tolist())
    # float64 items by bit pattern, parallel path
    f = cb.SpaceSaving(50, "f8")
    xf = rng.choice(np.array([0.0, -0.0, 1.5, np.nan, np.inf, -2.25]), 300000)
    f.update(xf)
    of = R.SpaceSaving(50)
    of.update(xf.view("i8"))
    assert f.counters().view("i8").tobytes() == of.counters().view("i8").tobytes()


def _assert_valid_summary(c, items, capacity, weights=None):
    """Metwally's guarantees against the true (weighted) counts."""
    if weights is None:
        vals, counts = np.unique(items, return_counts=True)
    else:
        vals, inv = np.unique(items, return_inverse=True)
        counts = np.bincount(inv, weights=weights).astype(np.int64)
    true = dict(zip(vals.tolist(), counts.tolist()))
    assert (np.diff(c["count"]) <= 0).all()
    for it, cn, er in c.tolist():
        assert cn - er <= true.get(it, 0) <= cn, (it, cn, er, true.get(it, 0))
    present = set(c["item"].tolist())
    if len(c) < capacity:
        assert present == set(vals.tolist())
    else:
        floor = c["count"].min()
        assert set(vals[counts > floor].tolist()) <= present
    return true


@pytest.mark.parametrize("n", [50000, 5 * 10**6])
def test_spsv_counting_path_exact_regime(cb, n):
    """Candidate counting (mode 3; auto from 65536 elements): with #distinct <= capacity the result
    is the reference's, bit for bit, list order included -- rare items that the sample missed come
    back through the ordered replay of the misses."""
    rng = np.random.default_rng(n)
    items = (rng.zipf(1.3, n) % 700).astype(np.int64) * 7919 - 3
    items[rng.integers(0, n, 5)] = np.iinfo(np.int64).min        # the lookup's empty marker
    items[[n // 3, n // 2]] = [123456789, -987654321]            # two singletons
    s, o = cb.SpaceSaving(1000, "i8"), R.SpaceSaving(1000)
    s.update(cb.DeviceArray.from_numpy(items), mode=0 if n >= 1 << 16 else 3)
    o.update(items)
    assert s.counters().tobytes() == o.counters().tobytes()
    s.update(items[::-1].copy(), mode=3)                         # on top of an existing summary
    o.update(items[::-1].copy())
    assert s.counters().tobytes() == o.counters().tobytes()
    assert (s.counters()["error"] == 0).all()


@pytest.mark.parametrize("kind", ["zipf", "uniform", "two_level"])
def test_spsv_counting_path_is_a_valid_summary(cb, kind):
    """More distinct items than counters: exact counts for the kept items, and every guarantee of
    a Space-Saving summary holds against the true counts (heavy tail, no heavy hitters at all,
    and a few heavy items over a sea of singletons)."""
    rng = np.random.default_rng(11)
    n = 3 * 10**6
    if kind == "zipf":
        items = rng.zipf(1.1, n).astype(np.int64)
    elif kind == "uniform":
        items = rng.integers(0, 1 << 40, n).astype(np.int64)
    else:
        items = np.where(rng.random(n) < 0.7, rng.integers(0, 300, n), rng.integers(1 << 20, 1 << 50, n)).astype(np.int64)
    for cap in (1000, 20):
        s = cb.SpaceSaving(cap, "i8")
        s.update(cb.DeviceArray.from_numpy(items), mode=3)
        c = s.counters()
        true = _assert_valid_summary(c, items, cap)
        if kind != "uniform":
            vals = np.array(list(true.keys())); counts = np.array(list(true.values()))
            k = min(cap // 2, 100)
            top = vals[np.argsort(-counts, kind="stable")[:k]]
            got = {it: (cn, er) for it, cn, er in c.tolist()}
            for it in top.tolist():                               # heavy hitters: exact counts
                assert got[it][0] - got[it][1] == true[it]
        # a second batch, then a merge with a replayed summary: still valid for the union
        s.update(items[: n // 2].copy(), mode=3)
        _assert_valid_summary(s.counters(), np.concatenate([items, items[: n // 2]]), cap)


@pytest.mark.parametrize("dtype,n", [("i8", 3 * 10**6), ("f8", 400000), ("i8", 30000)])
def test_spsv_and_moments_in_one_pass(cb, dtype, n):
    """BASELINE config 5: SummaryStats moments + SpaceSaving over the same items in ONE pass
    (crick_spsv_update_stats): the summary is the one the plain update builds, byte for byte, and
    the moments are those of the stand-alone SummaryStats pass (count / min / max exactly, the
    rest within the 1e-12 class of the parallel pass; test_stats.py:20-21 style check vs NumPy)."""
    rng = np.random.default_rng(n)
    items = rng.zipf(1.2, n).astype(np.int64)
    if dtype == "f8":
        items = (items % 5000).astype(np.float64) * 0.25
        items[::1000] = np.nan
    for dev in (True, False):
        s1, st1 = cb.SpaceSaving(1000, dtype), cb.SummaryStats()
        s2, st2 = cb.SpaceSaving(1000, dtype), cb.SummaryStats()
        x = cb.DeviceArray.from_numpy(items) if dev else items
        s1.update(x, stats=st1)
        s2.update(x)
        st2.update(x)
        assert s1.counters().tobytes() == s2.counters().tobytes()
        a, b = st1.__getstate__(), st2.__getstate__()
        assert a[0] == b[0] == np.count_nonzero(~np.isnan(items.astype(np.float64)))
        assert a[2] == b[2] and a[3] == b[3] and a[7] == b[7] and a[8] == b[8]
        for u, v in zip(a[1:7], b[1:7]):
            assert abs(u - v) <= 1e-11 * abs(v)
        xs = items.astype(np.float64)
        assert abs(st1.mean() - np.nanmean(xs)) <= 1e-12 * abs(np.nanmean(xs))
        assert abs(st1.var() - np.nanvar(xs)) <= 1e-9 * np.nanvar(xs)
    # a second batch on top (the summary is no longer fresh), stats accumulate
    s1.update(x, stats=st1)
    s2.update(x)
    assert s1.counters().tobytes() == s2.counters().tobytes()
    assert st1.count() == 2 * a[0]
    with pytest.raises(ValueError):
        s1.update(x, 2, stats=st1)


def test_spsv_records_merge_like_the_reference(cb):
    """Multi-GPU tail: records exported on the device, concatenated as an
    all-gather would deliver them, and merged in rank order must equal
    SpaceSaving(cap).merge(rank0, rank1, ...) of the reference, list order included -- full and
    non-full summaries, overlapping and disjoint items."""
    rng = np.random.default_rng(21)
    for cap, lens in ((50, (4000, 30, 2500, 0)), (1000, (300000, 200000))):
        parts, refs = [], []
        for k, n in enumerate(lens):
            x = (rng.zipf(1.2, n) % (400 if k % 2 else 3000)).astype(np.int64) + 17 * k
            s, o = cb.SpaceSaving(cap, "i8"), R.SpaceSaving(cap)
            if n:
                s.update(x, mode=1)
                o.update(x)
            parts.append(s)
            refs.append(o)
        nd = parts[0].record_int64s()
        assert nd == 2 + 3 * cap
        buf = cb.DeviceArray((len(parts), nd), "i8")
        for k, s in enumerate(parts):
            s.export_record(buf.ptr + 8 * nd * k)
        out, ref = cb.SpaceSaving(cap, "i8"), R.SpaceSaving(cap)
        out.merge_records(buf.ptr, len(parts))
        ref.merge(*refs)
        assert out.counters().tobytes() == ref.counters().tobytes(), cap
        rec0 = buf.to_numpy()[0]
        assert rec0[0] == len(refs[0].counters())
        assert np.array_equal(rec0[2:2 + 3 * int(rec0[0])].reshape(-1, 3),
                              np.stack([refs[0].counters()[f] for f in ("item", "count", "error")], 1))


def test_spsv_bulk_merge_of_records_and_device_stats_records(cb):
    """Multi-GPU tail on one GPU: 8 shard summaries -> 8 records -> merged (a) one after the other
    (the reference's spsv_merge, bit for bit) and (b) by the k-way form in one kernel.  (b) must be
    a valid summary of the union with the same guarantees, and when every distinct item fits the
    capacity both hold the exact counts.  SummaryStats records merged on the device == merge()."""
    rng = np.random.default_rng(77)
    n, cap = 1600000, 1000
    for distinct in (600, 50000):
        items = (rng.zipf(1.2, n) % distinct).astype(np.int64) * 977 + 5
        parts, sts = [], []
        nd = cb.SpaceSaving(cap, "i8").record_int64s()
        recs = cb.DeviceArray((8, nd), "i8")
        srec = cb.DeviceArray((8, 9))
        for k in range(8):
            sl = items[k * n // 8:(k + 1) * n // 8]
            p, st = cb.SpaceSaving(cap, "i8"), cb.SummaryStats()
            p.update(sl, stats=st)
            p.export_record(recs.ptr + 8 * nd * k)
            st.export_record_device(srec.ptr + 72 * k)
            parts.append(p); sts.append(st)
        from crick_b200._lib import lib
        lib.crick_device_synchronize()
        exact, bulk = cb.SpaceSaving(cap, "i8"), cb.SpaceSaving(cap, "i8")
        exact.merge_records(recs.ptr, 8)
        bulk.merge_records(recs.ptr, 8, bulk=True)
        via = cb.SpaceSaving(cap, "i8")
        via.merge(*parts)
        assert exact.counters().tobytes() == via.counters().tobytes()
        cb_ = bulk.counters()
        true = _assert_valid_summary(cb_, items, cap)
        if distinct <= cap:
            got = {it: (cn, er) for it, cn, er in cb_.tolist()}
            assert all(got[it] == (c, 0) for it, c in true.items())
            assert sorted(exact.counters().tolist()) == sorted(cb_.tolist())
        # on top of a summary that already holds data: still valid for the union
        bulk.merge_records(recs.ptr, 4, bulk=True)
        _assert_valid_summary(bulk.counters(), np.concatenate([items, items[: n // 2]]), cap)
        bulk.clear()
        assert bulk.size() == 0
        bulk.merge_records(recs.ptr, 8, bulk=True)
        assert bulk.counters().tobytes() == cb_.tobytes()
        # moments: device records merged in order == merge(*stats)
        a, b = cb.SummaryStats(), cb.SummaryStats()
        a.merge_records_device(srec.ptr, 8)
        b.merge(*sts)
        assert a.__getstate__() == b.__getstate__()
        a.clear()
        assert a.count() == 0 and np.isnan(a.mean())


def test_sketch_allreduce_world1_nccl(cb):
    """SketchAllReduce (one all-gather for summary + moments) on a 1-rank NCCL group."""
    import torch
    import torch.distributed as dist
    from crick_b200.distributed import SketchAllReduce
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29541", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        rng = np.random.default_rng(5)
        items = rng.zipf(1.3, 300000).astype(np.int64)
        sp, st = cb.SpaceSaving(100, "i8"), cb.SummaryStats()
        sp.update(items, stats=st)
        s = torch.cuda.Stream()
        for exact in (True, False):
            red = SketchAllReduce(100, "i8", exact=exact)
            with torch.cuda.stream(s):
                msp, mst = red(sp, st)
                msp, mst = red(sp, st)          # reused buffers and outputs
            s.synchronize()
            assert sorted(msp.counters().tolist()) == sorted(sp.counters().tolist())
            # (stats_merge :78-90 into an empty summary leaves first_value alone, like the reference)
            assert mst.__getstate__()[:8] == st.__getstate__()[:8]
    finally:
        if created:
            dist.destroy_process_group()


def test_spsv_allgather_world1_nccl(cb):
    import torch
    import torch.distributed as dist
    from crick_b200.distributed import allgather_merge_spsv
    rng = np.random.default_rng(5)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29618", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        x = rng.zipf(1.3, 100000).astype(np.int64)
        s, o = cb.SpaceSaving(100, "i8"), R.SpaceSaving(100)
        s.update(x, mode=1)
        o.update(x)
        m = allgather_merge_spsv(s)
        mo = R.SpaceSaving(100)
        mo.merge(o)
        assert m.counters().tobytes() == mo.counters().tobytes()
    finally:
        if created:
            dist.destroy_process_group()


def test_spsv_counting_path_degenerate_streams(cb):
    """One distinct item, two alternating items, and every element distinct but fewer than the
    capacity: the counting path (auto) must still give the reference's bits."""
    n = 1 << 22
    streams = [np.full(n, 7, dtype=np.int64),
               np.where(np.arange(n) % 3 == 0, -5, 11).astype(np.int64),
               np.concatenate([np.arange(900, dtype=np.int64), np.full(n - 900, 3, dtype=np.int64)])]
    for items in streams:
        s, o = cb.SpaceSaving(1000, "i8"), R.SpaceSaving(1000)
        s.update(cb.DeviceArray.from_numpy(items))
        o.update(items)
        assert s.counters().tobytes() == o.counters().tobytes()


@pytest.mark.parametrize("distinct", [300, 995, 1005, 20000])
def test_spsv_small_batches_keep_the_reference_bits(cb, distinct):
    """Below 65536 elements auto mode promises the reference's bits for every input.  A fresh
    summary may take the counting path, but only when every distinct item gets a counter (no
    eviction possible); otherwise -- and on a summary that already holds data -- the stream is
    replayed in order."""
    rng = np.random.default_rng(distinct)
    n = 30000
    items = (rng.zipf(1.2, n) % distinct).astype(np.int64) * 31 + 5
    s, o = cb.SpaceSaving(1000, "i8"), R.SpaceSaving(1000)
    s.update(cb.DeviceArray.from_numpy(items))
    o.update(items)
    assert s.counters().tobytes() == o.counters().tobytes()
    s.update(items[::-1].copy())
    o.update(items[::-1].copy())
    assert s.counters().tobytes() == o.counters().tobytes()


def test_spsv_counting_path_with_counts(cb):
    """A count per element (pre-aggregated input): candidate counts and miss buckets are sums of
    counts.  Exact regime == the reference; otherwise every guarantee holds for the weighted
    truth."""
    rng = np.random.default_rng(17)
    n = 400000
    w = rng.integers(1, 1000, n).astype(np.int64)
    items = (rng.zipf(1.25, n) % 800).astype(np.int64)
    s, o = cb.SpaceSaving(1000, "i8"), R.SpaceSaving(1000)
    s.update(cb.DeviceArray.from_numpy(items), cb.DeviceArray.from_numpy(w))
    o.update(items, w)
    assert s.counters().tobytes() == o.counters().tobytes()
    items = rng.zipf(1.1, n).astype(np.int64)
    for cap in (1000, 50):
        s = cb.SpaceSaving(cap, "i8")
        s.update(items, w, mode=3)
        _assert_valid_summary(s.counters(), items, cap, w)


@pytest.mark.gpu
def test_c_abi_nccl_comm_world1(cb):
    """crick_comm_* (NCCL opened by the library itself, no torch.distributed): a world of one rank
    reduces a TDigest, a SpaceSaving and a SummaryStats into fresh sketches that equal
    ``out.merge(sketch)``; the raw all-gather moves a device buffer.  (tests/tools/comm_check.py is
    the multi-process form; its 2-GPU output is under profiles/.)"""
    from crick_b200._lib import lib
    from crick_b200.distributed import NcclComm
    rng = np.random.default_rng(3)
    x = rng.normal(size=50000)
    items = rng.zipf(1.3, 50000).astype(np.int64)
    t = cb.TDigest(100.0); t.update(x)
    s = cb.SpaceSaving(32, "i8"); s.update(items)
    m = cb.SummaryStats(); m.update(x)
    comm = NcclComm(NcclComm.unique_id(), 0, 1)
    try:
        for exact in (True, False):
            rt = comm.reduce_tdigest(t, exact=exact)
            rs = comm.reduce_space_saving(s, exact=exact)
            rm = comm.reduce_stats(m)
            comm.synchronize()
            et = cb.TDigest(100.0); et.merge(t)
            es = cb.SpaceSaving(32, "i8"); es.merge(s)
            if exact:
                assert rt.centroids().tobytes() == et.centroids().tobytes()
            assert rt.size() == t.size()
            assert abs(rt.quantile(0.5) - t.quantile(0.5)) < 1e-3
            got, want = np.asarray(rs.counters()), np.asarray(es.counters())
            if exact:
                assert got.tobytes() == want.tobytes()
            else:                          # the k-way merge keeps the counters, not the order of ties
                rows = lambda a: sorted(a[i:i + 1].tobytes() for i in range(len(a)))
                assert rows(got) == rows(want)
            assert (rm.count(), rm.sum(), rm.var()) == (m.count(), m.sum(), m.var())
        a = cb.DeviceArray.from_numpy(np.arange(16.0)); b = cb.DeviceArray((16,))
        assert lib.crick_comm_allgather(comm._h, a.ptr, b.ptr, 128, comm.stream) == 0
        comm.synchronize()
        assert np.array_equal(b.to_numpy(), np.arange(16.0))
    finally:
        comm.close()
