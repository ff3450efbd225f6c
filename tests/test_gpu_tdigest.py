"""GPU parity tests (reference-compatible mode): the CUDA path, called through the C ABI of
libcrick_cuda.so by the drop-in wrappers, against (a) the committed golden vectors made from
the built reference (tests/golden/make_golden.py) and (b) the oracle on the same seeded inputs.
The bar is BIT-EXACT centroids / min / max / size and bit-exact quantile / cdf answers.

Edge cases follow the reference's own tests (crick/tests/test_tdigest.py): empty digest -> NaN
:129-136, single value :139-152, NaN/inf silently dropped :155-180, tiny weights dropped
:183-193, output shapes :230-251, histogram :254-344, weights :347-361, pickle :364-374,
merge :377-408, scale :411-444.
"""
import pickle

import numpy as np
import pytest

import oracle
from helpers import (QS5, QS9, check_digest_against_golden, kat_d_parts, kat_inputs,
                     quantiles_to_q, q_to_x, sha)

pytestmark = pytest.mark.gpu
R = oracle.restated


@pytest.fixture(scope="module")
def cb():
    import crick_b200
    if crick_b200.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return crick_b200


def same(a, b):
    """Digest `a` (ours) and `b` (oracle) are identical bit for bit."""
    assert a.centroids().tobytes() == b.centroids().tobytes()
    assert a.size() == b.size()
    assert np.array_equal([a.min(), a.max()], [b.min(), b.max()], equal_nan=True)


# ---- golden vectors from the built reference -----------------------------------------
@pytest.mark.parametrize("name", ["kat_a", "kat_b", "kat_c"])
def test_kats_bit_exact(cb, golden, name):
    g, _ = golden
    c, x, w, cdf_x = kat_inputs()[name]
    t = cb.TDigest(c, mode="reference")
    t.update(x, w)
    check_digest_against_golden(t, g[name], cdf_x)


def test_kat_d_merge_bit_exact(cb, golden):
    g, _ = golden
    parts = []
    for x in kat_d_parts():
        p = cb.TDigest(100, mode="reference")
        p.update(x)
        parts.append(p)
    t = cb.TDigest(100)
    t.merge(*parts)
    check_digest_against_golden(t, g["kat_d"], [0.5, 4, 7.5])


def test_config1_bit_exact(cb, golden):
    """BASELINE.json configs[0]: TDigest(100).update of 1e6 uniform then 5 quantiles."""
    g, _ = golden
    t = cb.TDigest(100, mode="reference")
    t.update(np.random.default_rng(0).uniform(0, 1, 10**6))
    check_digest_against_golden(t, g["config1"], [0.1, 0.5, 0.9])


def test_config2_small_bit_exact(cb, golden):
    g, _ = golden
    rng = np.random.default_rng(1234)
    x = rng.lognormal(0, 1, 200000)
    w = rng.uniform(0.5, 1.5, 200000)
    t = cb.TDigest(100, mode="reference")
    t.update(x, w)
    check_digest_against_golden(t, g["config2_small"], [0.5, 1.0, 3.0])


# ---- oracle on the same seeded inputs --------------------------------------------------
@pytest.mark.parametrize("c", [20, 50.5, 100, 200, 500, 1000])
def test_weighted_stream_vs_oracle(cb, c):
    rng = np.random.default_rng(int(c))
    x = rng.lognormal(size=30000)
    w = rng.uniform(0.5, 1.5, 30000)
    a = cb.TDigest(c, mode="reference")
    b = R.TDigest(c)
    a.update(x, w)
    b.update(x, w)
    same(a, b)
    qs = np.concatenate([rng.uniform(-0.1, 1.1, 1000), [0.0, 1.0, 1e-12, 1 - 1e-12]])
    assert a.quantile(qs).tobytes() == b.quantile(qs).tobytes()
    cm = b.centroids()["mean"]
    xs = np.concatenate([rng.lognormal(size=1000), cm, np.nextafter(cm, np.inf),
                         np.nextafter(cm, -np.inf), [b.min(), b.max(), -1.0, 1e9]])
    assert a.cdf(xs).tobytes() == b.cdf(xs).tobytes()


@pytest.mark.parametrize("kind", ["sorted", "reversed", "ties", "zeros", "step", "narrow"])
def test_hard_orders_vs_oracle(cb, kind):
    rng = np.random.default_rng(7)
    n = 20000
    x = {"sorted": np.sort(rng.normal(size=n)),
         "reversed": np.sort(rng.normal(size=n))[::-1].copy(),
         "ties": rng.integers(0, 13, n).astype(float),
         "zeros": rng.choice([0.0, -0.0, 1.0, -1.0], n),
         "step": np.concatenate([rng.uniform(0, 1, n // 2), rng.uniform(100, 101, n // 2)]),
         "narrow": rng.normal(50, 1e-9, n)}[kind]
    a = cb.TDigest(100, mode="reference")
    b = R.TDigest(100)
    a.update(x)
    b.update(x)
    same(a, b)
    assert a.quantile(QS9).tobytes() == b.quantile(QS9).tobytes()


def test_buffer_phase_carries_across_calls(cb):
    """tdigest_add's flush-before-append phase (tdigest_stubs.c:288-290) must carry across
    update()/add()/merge() calls of any size, including sizes that straddle buffer_size."""
    rng = np.random.default_rng(11)
    a = cb.TDigest(100, mode="reference")
    b = R.TDigest(100)
    for n in [1, 41, 1, 1, 42, 43, 0, 7, 500, 84, 3]:
        x = rng.normal(size=n)
        a.update(x)
        b.update(x)
        assert a.size() == b.size()
    for v in rng.normal(size=50):
        a.add(v, 2.0)
        b.add(v, 2.0)
    same(a, b)


def test_nonfinite_dropped_and_tiny_weights(cb):
    x = np.array([1.0, np.nan, 2.0, np.inf, 3.0, -np.inf, 4.0])
    a = cb.TDigest(mode="reference")
    b = R.TDigest()
    a.update(x)
    b.update(x)
    same(a, b)
    assert a.size() == 4
    eps = np.finfo("f8").eps
    a = cb.TDigest(mode="reference")
    a.update([1.0, 2.0, 3.0], [eps, 1.0, eps / 2])  # w <= eps dropped (tdigest_stubs.c:285)
    assert a.size() == 1.0 and len(a.centroids()) == 1
    with pytest.raises(ValueError):
        a.update([1.0], [0.0])
    with pytest.raises(ValueError):
        a.update([1.0], [np.nan])
    with pytest.raises(TypeError):
        a.update(["a"])


def test_empty_single_and_shapes(cb):
    t = cb.TDigest()
    assert np.isnan(t.min()) and np.isnan(t.max()) and t.size() == 0
    assert np.isnan(t.quantile(0.5)) and np.isnan(t.cdf(1.0))
    assert len(t.centroids()) == 0
    t.update([])
    assert t.size() == 0
    t.add(10)
    assert [t.quantile(q) for q in (0, 0.5, 1)] == [10, 10, 10]
    assert [t.cdf(x) for x in (9, 10, 11)] == [0, 0.5, 1]
    assert isinstance(t.quantile(0.5), np.float64)
    assert t.quantile(np.zeros((2, 3))).shape == (2, 3)
    assert t.cdf(np.zeros((0,))).shape == (0,)
    assert repr(t) == "TDigest<compression=100.0, size=1.0>"
    assert cb.TDigest(5).compression == 20.0 and cb.TDigest(5000).compression == 1000.0


def test_korder_inputs(cb, golden):
    """Non-contiguous inputs are consumed in memory (K) order like NpyIter does
    (tdigest_stubs.c:652-666); shas recorded from the built reference."""
    g, a = golden
    base, wv = a["korder_base"], a["korder_w"]
    cases = {"rev": (base.ravel()[::-1], 1), "fortran": (np.asfortranarray(base), 1),
             "transposed": (base.T, 1), "strided": (base.ravel()[::2], 1),
             "rev_fw": (base.ravel()[::-1], wv.ravel()), "bcast_w": (base, wv[0]),
             "t_w": (base.T, wv.T), "scalar_x": (1.5, wv.ravel()[:100])}
    for name, (x, w) in cases.items():
        t = cb.TDigest(100, mode="reference")
        t.update(x, w)
        assert sha(t) == g["korder"][name], name


def test_merge_semantics_vs_oracle(cb):
    rng = np.random.default_rng(3)
    ours, theirs = [], []
    for k, c in enumerate([100, 100, 500, 50, 100]):
        x = rng.normal(k, 1 + k, 3000 + 17 * k)
        p, o = cb.TDigest(c, mode="reference"), R.TDigest(c)
        p.update(x)
        o.update(x)
        ours.append(p)
        theirs.append(o)
    a, b = cb.TDigest(100, mode="reference"), R.TDigest(100)
    # 10 values stay in the buffer: merge feeds through it and does not flush T (:592-606)
    x0 = rng.normal(size=10)
    a.update(x0)
    b.update(x0)
    a.merge(*ours)
    b.merge(*theirs)
    assert a.size() == b.size()
    same(a, b)
    for p, o in zip(ours, theirs):   # merge flushes the others but does not change them
        same(p, o)
    empty = cb.TDigest()
    a.merge(empty)
    same(a, b)
    with pytest.raises(TypeError):
        a.merge(1)


def test_scale_pickle_histogram(cb):
    rng = np.random.default_rng(5)
    x = rng.gamma(2.0, size=50000)
    a, b = cb.TDigest(mode="reference"), R.TDigest()
    a.update(x)
    b.update(x)
    for f in (0.5, 2, 1e-3, np.finfo("f8").eps):
        same(a.scale(f), b.scale(f))
    with pytest.raises(ValueError):
        a.scale(0)
    a2 = pickle.loads(pickle.dumps(a))
    same(a2, b)
    assert a2.compression == a.compression
    # quirk: compaction copies weight but not mean (tdigest_stubs.c:621)
    t = cb.TDigest()
    t.update([1, 2, 3, 4, 5], [1, 1000, 1, 10000, 1])
    c = t.scale(np.finfo("f8").eps).centroids()
    assert c["mean"].tolist() == [1.0, 2.0] and len(c) == 2
    h, e = a.histogram(20)
    cdf = b.cdf(e)
    assert np.array_equal(h, (cdf[1:] - cdf[:-1]) * b.size())
    assert abs(h.sum() - a.size()) < 1e-6 * a.size()
    h2, e2 = a.histogram([0.0, 1.0, 2.0, 100.0])
    assert h2.shape == (3,)
    with pytest.raises(ValueError):
        a.histogram([2.0, 1.0])


@pytest.mark.parametrize("dist", ["gamma", "uniform", "normal", "sequential", "step"])
def test_rank_error_bounds_of_the_reference_tests(cb, dist):
    """crick/tests/test_tdigest.py:72-107: monotone quantiles, rank error <= 0.012,
    cdf error <= 0.005, exact size/min/max -- on both update modes."""
    rng = np.random.default_rng(21)
    n = 100000
    data = {"gamma": rng.gamma(0.1, 0.1, n), "uniform": rng.uniform(0, 1, n),
            "normal": rng.normal(0, 1e-5, n), "sequential": np.arange(n) * 1e-5,
            "step": np.concatenate([rng.uniform(0, 1, n // 2), rng.uniform(2, 3, n // 2)])}[dist]
    for mode in ("reference", "parallel"):
        t = cb.TDigest(mode=mode)
        t.update(data)
        assert t.size() == n and t.min() == data.min() and t.max() == data.max()
        qs = np.linspace(0.001, 0.999, 100)
        est = t.quantile(qs)
        assert (np.diff(est) > 0).all()
        c = t.cdf(est)
        assert ((c >= 0) & (c <= 1)).all()
        assert np.abs(quantiles_to_q(data, t.quantile(QS9)) - QS9).max() <= 0.012, mode
        assert np.abs(t.cdf(q_to_x(data, QS9)) - QS9).max() <= 0.005, mode


def test_device_array_inputs_match_host_inputs(cb):
    rng = np.random.default_rng(9)
    x = rng.lognormal(size=70000)
    w = rng.uniform(0.5, 1.5, 70000)
    a, b = cb.TDigest(mode="reference"), cb.TDigest(mode="reference")
    a.update(x, w)
    b.update(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w))
    assert a.centroids().tobytes() == b.centroids().tobytes()
    qd = cb.DeviceArray.from_numpy(QS9)
    assert a.quantile(qd).to_numpy().tobytes() == a.quantile(QS9).tobytes()
    import torch
    tx = torch.from_numpy(x).cuda()
    c = cb.TDigest(mode="reference")
    c.update(tx, 2.0)           # __cuda_array_interface__ / DLPack of a foreign library
    d = R.TDigest()
    d.update(x, 2.0)
    assert c.centroids().tobytes() == d.centroids().tobytes()


# ---- grouped digests and the merge tree (BASELINE configs[2], [3]) ---------------------
@pytest.fixture(params=[0, 1], ids=["warp_per_digest", "thread_per_digest"])
def group_kernel(request):
    """Both grouped ingest kernels (tdigest_kernels.cu k_ingest, tdigest_lanes.cu) must give the
    reference's bits; auto would pick by group count."""
    from crick_b200._lib import lib
    prev = lib.crick_group_kernel(request.param)
    yield request.param
    lib.crick_group_kernel(prev)


def test_grouped_fixed_rows_vs_oracle(cb, group_kernel):
    rng = np.random.default_rng(2)
    vals = rng.normal(size=(3000, 1000))
    g = cb.TDigestGroup(3000, 100)
    g.update(vals)
    rows = list(range(0, 3000, 53))
    for r in rows:
        b = R.TDigest(100)
        b.update(vals[r])
        assert g.centroids(r).tobytes() == b.centroids().tobytes(), r
    q = g.quantile([0.1, 0.5, 0.9])
    c = g.cdf([-1.0, 0.0, 1.0])
    for r in rows[:10]:
        b = R.TDigest(100)
        b.update(vals[r])
        assert q[r].tobytes() == b.quantile([0.1, 0.5, 0.9]).tobytes()
        assert c[r].tobytes() == b.cdf([-1.0, 0.0, 1.0]).tobytes()
    meta = g.meta()
    assert (meta[:, 1] + 0 >= 0).all()


def test_grouped_ragged_weighted_and_incremental(cb, group_kernel):
    rng = np.random.default_rng(4)
    lens = rng.integers(0, 400, 500)
    lens[7] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    v = rng.lognormal(size=off[-1])
    w = rng.uniform(0.5, 2.0, off[-1])
    g = cb.TDigestGroup(500, 200)
    g.update(v, w, offsets=off)
    g.update(v, w, offsets=off)          # second batch continues each group's buffer phase
    for r in range(0, 500, 11):
        b = R.TDigest(200)
        sl = slice(off[r], off[r + 1])
        if lens[r]:
            b.update(v[sl], w[sl])
            b.update(v[sl], w[sl])
        assert g.centroids(r).tobytes() == b.centroids().tobytes(), r
    d = g.digest(3)
    assert d.centroids().tobytes() == g.centroids(3).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("margin", [6e-13, 1e-3, 0.2, 2.0, -1.0])
def test_group_decision_window_is_exact_for_any_margin(cb, margin):
    """The thread-per-digest kernel decides from a weight window and falls back to the reference's
    kscale pair inside it (tdigest_lanes.cu).  Whatever the window -- the certified default, wide
    ones that send a few percent / most / all items through the fallback, or the kernel without
    windows (-1) -- every byte equals the warp-per-digest kernel (which evaluates every kscale),
    and the fallback count moves as the window does."""
    import ctypes
    from crick_b200._lib import lib
    rng = np.random.default_rng(5)
    ng, row = 1500, 700
    v = np.concatenate([rng.normal(size=(ng // 2, row)), rng.lognormal(0, 2, size=(ng - ng // 2, row))])
    w = rng.uniform(0.5, 2.0, v.shape)
    out = {}
    nfb = ctypes.c_int64(0)
    for mode, mg in ((0, 6e-13), (1, margin)):
        prev = lib.crick_group_kernel(mode)
        pm = lib.crick_group_margin(mg)
        try:
            lib.crick_group_exact_fallbacks(ctypes.byref(nfb), 1)
            g = cb.TDigestGroup(ng, 100)
            g.update(v, w)
            g.update(v[:, ::-1].copy())            # unweighted second pass over existing centroids
            g.flush()
            lib.crick_group_exact_fallbacks(ctypes.byref(nfb), 1)
            out[mode] = (g.meta().copy(), [g.centroids(r) for r in range(0, ng, 7)], nfb.value)
        finally:
            lib.crick_group_kernel(prev)
            lib.crick_group_margin(pm)
    assert out[0][0].tobytes() == out[1][0].tobytes()
    for a, b in zip(out[0][1], out[1][1]):
        assert a.tobytes() == b.tobytes()
    assert out[0][2] == 0                           # the warp-per-digest kernel never counts
    n_items = 2 * v.size
    if margin == 6e-13:
        assert out[1][2] < 1e-4 * n_items
    elif margin == 1e-3:
        assert 0 < out[1][2] < 2.0 * n_items
    elif margin >= 0.2:
        assert out[1][2] > 0.5 * n_items
    else:
        assert out[1][2] == 0


@pytest.mark.parametrize("comp", [20, 100, 333.5, 1000])
def test_group_kernels_agree_on_hostile_input(cb, comp):
    """Thread-per-digest vs warp-per-digest on the same input: NaN / inf (dropped, :285), +-0.0
    ties, duplicates, tiny weights, ragged lengths incl. empty groups, pending buffers
    carried over a second update -- every byte of the state must agree, and a sample of groups
    must equal the reference."""
    from crick_b200._lib import lib
    rng = np.random.default_rng(int(comp))
    ng = 700
    lens = rng.integers(0, 900, ng)
    lens[[3, 77]] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(off[-1])
    v = rng.normal(size=n).round(1)                       # many duplicates
    v[rng.random(n) < 0.02] = np.nan
    v[rng.random(n) < 0.01] = np.inf
    v[rng.random(n) < 0.02] = -0.0
    v[rng.random(n) < 0.02] = 0.0
    w = rng.uniform(0.1, 3.0, n)
    w[rng.random(n) < 0.02] = 1e-17                        # <= eps: dropped
    states = []
    for mode in (0, 1):
        prev = lib.crick_group_kernel(mode)
        try:
            g = cb.TDigestGroup(ng, comp)
            g.update(v, w, offsets=off)
            g.update(v[::-1].copy(), 2.0, offsets=(n - off[::-1]).astype(np.int64))
            meta_pending = g.meta().copy()
            cents = [g.centroids(r) for r in range(ng)]    # flushes
            states.append((meta_pending, g.meta().copy(), cents))
        finally:
            lib.crick_group_kernel(prev)
    a, b = states
    assert a[0].tobytes() == b[0].tobytes()
    assert a[1].tobytes() == b[1].tobytes()
    for r in range(ng):
        assert a[2][r].tobytes() == b[2][r].tobytes(), r
    # reference check of the second-update composition on a few groups
    offr = (n - off[::-1]).astype(np.int64)
    vr = v[::-1].copy()
    for r in range(0, ng, 41):
        o = R.TDigest(comp)
        s1 = slice(off[r], off[r + 1]); s2 = slice(offr[r], offr[r + 1])
        if s1.stop > s1.start:
            o.update(v[s1], w[s1])
        if s2.stop > s2.start:
            o.update(vr[s2], 2.0)
        assert a[2][r].tobytes() == o.centroids().tobytes(), r


@pytest.mark.parametrize("ngroups", [1, 2, 37, 64])
def test_tree_merge_vs_nested_oracle_merges(cb, ngroups):
    rng = np.random.default_rng(ngroups)
    v = rng.normal(size=(ngroups, 2000))
    g = cb.TDigestGroup(ngroups, 200)
    g.update(v)
    tm = g.merge_tree()
    ds = []
    for r in range(ngroups):
        b = R.TDigest(200)
        b.update(v[r])
        ds.append(b)
    s = 1
    while s < ngroups:
        for k in range(0, ngroups, 2 * s):
            if k + s < ngroups:
                ds[k].merge(ds[k + s])
        s *= 2
    same(tm, ds[0])
    qs = rng.uniform(0, 1, 5000)
    assert tm.quantile(qs).tobytes() == ds[0].quantile(qs).tobytes()
    xs = rng.normal(size=5000)
    assert tm.cdf(xs).tobytes() == ds[0].cdf(xs).tobytes()


@pytest.mark.parametrize("ngroups,per,comp", [(3, 50, 100), (700, 1500, 200), (4096, 300, 100)])
def test_merge_all_one_pass_vs_the_exact_tree(cb, ngroups, per, comp):
    """TDigestGroup.merge_all: every centroid of every digest as a (mean, weight) point through
    one parallel k-bucket update.  Not the reference's bits (merge_tree has those); checked like
    SURVEY.md 8c prescribes for the parallel mode: exact min / max, total weight, and a rank error
    at the 9 standard q's no worse than the bit-exact tree's (plus the reference tests' caps,
    test_tdigest.py:399-405).  The group is left usable (flushed, not consumed)."""
    rng = np.random.default_rng(ngroups)
    v = rng.lognormal(size=(ngroups, per)) * np.linspace(0.5, 2.0, ngroups)[:, None]
    g = cb.TDigestGroup(ngroups, comp)
    g.update(v)
    fast = g.merge_all()
    again = g.merge_all()
    assert fast.centroids().tobytes() == again.centroids().tobytes()     # deterministic, group intact
    # merging INTO a digest that already holds data keeps it
    t = cb.TDigest(comp)
    t.update(np.full(1000, -5.0))
    from crick_b200._lib import lib, check
    check(lib.crick_tdigest_group_merge_all(g._h, t._h))
    assert t.min() == -5.0 and t.max() == v.max() and abs(t.size() - (v.size + 1000)) <= 1e-9 * v.size
    tree = g.merge_tree()                                                # (consumes the group)
    assert fast.min() == v.min() == tree.min() and fast.max() == v.max() == tree.max()
    assert abs(fast.size() - v.size) <= 1e-9 * v.size
    assert len(fast.centroids()) <= 2 * int(np.ceil(comp))
    flat = np.sort(v.ravel())
    rank = lambda q: np.searchsorted(flat, q) / flat.size
    ef = np.abs(rank(fast.quantile(QS9)) - QS9).max()
    et = np.abs(rank(tree.quantile(QS9)) - QS9).max()
    assert ef <= max(1.5 * et, 2e-3 if ngroups > 3 else 0.012), (ef, et)
    assert (np.diff(fast.quantile(np.linspace(0, 1, 101))) >= 0).all()


def test_cuda_array_inputs_of_other_libraries_and_dtypes(cb):
    """tdigest.pyx:298-305 for CUDA arrays: every dtype that casts safely to float64 is accepted
    (cast on the device), device weights are validated in reference mode too, inputs produced
    asynchronously on another stream are waited for, and the outputs travel by DLPack."""
    import torch
    rng = np.random.default_rng(3)
    x = rng.normal(size=1500)
    # (a) dtypes: float32 / float16 / int64 / int32 / uint8 / bool tensors == their host versions
    for dt in (np.float32, np.float16, np.int64, np.int32, np.uint8, np.bool_):
        h = (x * 10).astype(dt)
        t1, t2 = cb.TDigest(100, mode="reference"), cb.TDigest(100, mode="reference")
        t1.update(torch.from_numpy(h).cuda())
        t2.update(h)
        assert t1.centroids().tobytes() == t2.centroids().tobytes(), dt
    big = rng.lognormal(size=300000).astype(np.float32)
    t1, t2 = cb.TDigest(100), cb.TDigest(100)
    t1.update(torch.from_numpy(big).cuda(), torch.ones(big.size, dtype=torch.float32, device="cuda"))
    t2.update(big.astype(np.float64), np.ones(big.size))
    assert t1.centroids().tobytes() == t2.centroids().tobytes()
    with pytest.raises(TypeError):
        cb.TDigest().update(torch.zeros(10, dtype=torch.complex64, device="cuda"))
    # (b) device weights in reference mode are checked like host weights
    t = cb.TDigest(100, mode="reference")
    xd = torch.from_numpy(x).cuda()
    for bad in (0.0, -1.0, float("nan"), float("inf")):
        w = torch.ones(x.size, dtype=torch.float64, device="cuda")
        w[7] = bad
        with pytest.raises(ValueError):
            t.update(xd, w)
    assert t.size() == 0
    t.update(xd, torch.full((x.size,), 2.0, dtype=torch.float64, device="cuda"))
    assert t.size() == 2 * x.size
    # (c) producer on a side stream, no stream= passed: the update must see the finished data
    side = torch.cuda.Stream()
    ref = cb.TDigest(100)
    for rep in range(20):
        with torch.cuda.stream(side):
            a = torch.empty(1 << 20, dtype=torch.float64, device="cuda")
            a.normal_(generator=None)
            a.mul_(3.0).add_(float(rep))
        # without waiting for `side` here: DLPack hands the ordering to the library
        d = cb.TDigest(100)
        with torch.cuda.stream(side):
            d.update(a)
        host = a.cpu().numpy()
        assert d.min() == host.min() and d.max() == host.max() and d.size() == host.size
        with torch.cuda.stream(side):
            d2 = cb.TDigest(100)
            d2.update(a, stream=side.cuda_stream)
        side.synchronize()
        assert d2.centroids().tobytes() == d.centroids().tobytes()
    # (d) DLPack export: DeviceArray and the zero-copy centroid view
    arr = cb.DeviceArray.from_numpy(x)
    tt = torch.from_dlpack(arr)
    assert tt.is_cuda and tt.dtype == torch.float64 and np.array_equal(tt.cpu().numpy(), x)
    del arr
    assert np.array_equal(tt.cpu().numpy(), x)              # the capsule keeps the buffer alive
    d = cb.TDigest(100)
    d.update(x)
    cv = torch.from_dlpack(d.centroids_device())
    c = d.centroids()
    assert cv.shape == (len(c), 2) and np.array_equal(cv.cpu().numpy(), c.view(np.float64).reshape(-1, 2))
    q = d.quantile(torch.tensor([0.1, 0.5, 0.9], dtype=torch.float32, device="cuda"))   # float32 queries
    assert np.allclose(torch.from_dlpack(q).cpu().numpy(), d.quantile(np.array([0.1, 0.5, 0.9], dtype=np.float32)))


def test_large_query_batch_vs_oracle(cb):
    """Config-4 query shape: millions of device-resident queries, bit-exact vs the oracle, plus
    the size-independent properties (monotone, cdf(quantile(q)) ~ q)."""
    rng = np.random.default_rng(8)
    x = rng.normal(size=200000)
    t, o = cb.TDigest(200, mode="reference"), R.TDigest(200)
    t.update(x)
    o.update(x)
    q = np.sort(rng.uniform(0.0005, 0.9995, 2 * 10**6))
    dq = cb.DeviceArray.from_numpy(q)
    dx = t.quantile(dq)
    xs = dx.to_numpy()
    assert xs.tobytes() == o.quantile(q).tobytes()
    back = t.cdf(dx).to_numpy()
    assert back.tobytes() == o.cdf(xs).tobytes()
    assert (np.diff(xs) >= 0).all()
    assert np.abs(back - q).max() < 1e-3


def test_allgather_reducer_world1_nccl(cb):
    """The multi-GPU reduction (export record -> NCCL all-gather -> rank-order merge kernel) on a
    1-rank group: must equal TDigest(c).merge(digest); the reducer's buffers are reused."""
    import torch
    import torch.distributed as dist
    from crick_b200.distributed import TDigestAllReduce, allgather_merge_tdigest
    rng = np.random.default_rng(13)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29617", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        red = TDigestAllReduce(100)
        s = torch.cuda.Stream()
        for _ in range(3):
            x = rng.lognormal(size=50000)
            t, o = cb.TDigest(100, mode="reference"), R.TDigest(100)
            t.update(x)
            o.update(x)
            with torch.cuda.stream(s):
                m = red(t)
            mo = R.TDigest(100)
            mo.merge(o)
            same(m, mo)
        m2 = allgather_merge_tdigest(t)
        same(m2, mo)
        t.clear()
        assert t.size() == 0 and len(t.centroids()) == 0 and np.isnan(t.min())
    finally:
        if created:
            dist.destroy_process_group()


def test_pickles_of_the_reference_load_and_round_trip(cb, golden):
    """sketches move between stock crick and this build.  The golden pickles
    were written by the reference itself (tests/golden/make_golden.py); they name
    crick.tdigest.TDigest / crick.stats.SummaryStats / crick.space_saving.SpaceSaving, which the
    `crick/` shim resolves to our classes.  Loading must give the reference's state, and
    re-pickling at the same protocol must give the reference's bytes."""
    import crick  # the shim package at the repository root
    g, _ = golden
    i = np.arange(10000)
    x = ((i * 7919) % 10007) / 10007.0
    t = pickle.loads(bytes.fromhex(g["pickle_tdigest"]))
    assert type(t) is cb.TDigest and type(t) is crick.TDigest
    o = R.TDigest(100)
    o.update(x)
    same(t, o)
    assert pickle.dumps(t, protocol=2) == bytes.fromhex(g["pickle_tdigest"])
    t.update(x[:100])                                   # a restored digest keeps working
    o.update(x[:100])
    same(t, o)
    s = pickle.loads(bytes.fromhex(g["pickle_stats"]))
    so = R.SummaryStats()
    so.update(x)
    assert type(s) is cb.SummaryStats and s.__getstate__() == so.__getstate__()
    assert pickle.dumps(s, protocol=2) == bytes.fromhex(g["pickle_stats"])
    for key, items in (("pickle_spsv_i8", [1, 1, 2, 3, 3, 3]), ("pickle_spsv_f8", [1.5, 1.5, 2.0])):
        sp = pickle.loads(bytes.fromhex(g[key]))
        ref = cb.SpaceSaving(5, key[-2:])
        ref.update(items)
        assert type(sp) is cb.SpaceSaving and sp.counters().tobytes() == ref.counters().tobytes()
        assert pickle.dumps(sp, protocol=2) == bytes.fromhex(g[key])


def test_dlpack_only_inputs_and_zero_copy_centroids(cb):
    """CUDA arrays may arrive as DLPack producers without __cuda_array_interface__, and
    centroids_device() exposes the live centroid table without a copy."""
    import torch
    rng = np.random.default_rng(31)
    x = rng.normal(size=100000)

    class DLPackOnly:
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, stream=None):
            return self._t.__dlpack__()

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    tx = torch.from_numpy(x).cuda()
    a, b = cb.TDigest(mode="reference"), R.TDigest()
    a.update(DLPackOnly(tx))
    b.update(x)
    same(a, b)
    q = torch.tensor([0.1, 0.5, 0.9], dtype=torch.float64, device="cuda")
    got = a.quantile(DLPackOnly(q)).to_numpy()
    assert got.tobytes() == b.quantile([0.1, 0.5, 0.9]).tobytes()
    a32, b32 = cb.TDigest(mode="reference"), R.TDigest()
    a32.update(DLPackOnly(tx.float()))                    # safely castable dtypes are cast on the device
    b32.update(x.astype(np.float32))
    same(a32, b32)
    with pytest.raises(TypeError):
        a.update(DLPackOnly(torch.complex(tx, tx)))       # tdigest.pyx:298-302: not castable to f8
    with pytest.raises(ValueError):
        a.update(DLPackOnly(torch.from_numpy(x.reshape(100, 1000)).cuda()[:, ::2]))   # strided
    view = a.centroids_device()
    cen = torch.as_tensor(view, device="cuda").cpu().numpy()
    c = a.centroids()
    assert cen.shape == (len(c), 2)
    assert np.array_equal(cen[:, 0], c["mean"]) and np.array_equal(cen[:, 1], c["weight"])
    # CPU-side DLPack producers are not device arrays: they take the NumPy path
    a2 = cb.TDigest(mode="reference")
    a2.update(torch.from_numpy(x))
    same(a2, b)
