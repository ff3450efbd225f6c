"""GPU parity tests, parallel (k-bucket) mode -- the path BASELINE.json's metric is quoted on.

The clustering differs from the reference's sequential one by construction, so parity is
checked the way SURVEY.md 8c prescribes:
  (a) the kernel's partition is recomputed on the CPU from the kernel's own bucket edges:
      per-bucket (sum w, sum w*x) within 1e-12 relative, exact min / max, and the digest equals
      the oracle's greedy merge of those bucket points BIT FOR BIT;
  (b) weighted rank error at the 9 standard q's <= the reference's on the same data (plus the
      absolute caps of crick/tests/test_tdigest.py:99-107);
  (c) size-independent properties at full size (1e9 values on the device): exact total weight
      by linearity, exact min/max, monotone quantiles.
"""
import pickle
import numpy as np
import pytest

import oracle
from helpers import QS9, accepted, restate_parallel_items, weighted_rank_fn

pytestmark = pytest.mark.gpu
R = oracle.restated


@pytest.fixture(scope="module")
def cb():
    import crick_b200
    if crick_b200.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return crick_b200


def check_partition(cb, x, w, comp=100.0, device=False):
    """(a): run one parallel update, then restate it on the CPU from the kernel's edges."""
    t = cb.TDigest(comp, mode="parallel")
    if device:
        t.update(cb.DeviceArray.from_numpy(x),
                 cb.DeviceArray.from_numpy(w) if np.ndim(w) else float(w))
    else:
        t.update(x, w)
    edges = t._debug_edges()
    bw, bs, mm = t._debug_buckets()
    xa, wa = accepted(x, w)
    assert (np.diff(edges) > 0).all()
    assert len(bw) == len(edges) + 1
    b = np.searchsorted(edges, xa, side="right")          # bucket = #edges <= x
    cw = np.bincount(b, weights=wa, minlength=len(bw))
    cs = np.bincount(b, weights=wa * xa, minlength=len(bw))
    assert np.allclose(bw, cw, rtol=1e-12, atol=0)
    scale = np.bincount(b, weights=np.abs(wa * xa), minlength=len(bw))
    assert (np.abs(bs - cs) <= 1e-12 * np.maximum(scale, 1e-300)).all()
    assert mm[0] == xa.min() and mm[1] == xa.max()
    # finalize: bucket points -> the reference's greedy merge (tdigest_flush phases 2-5)
    items = restate_parallel_items(edges, bw, bs, mm[0], mm[1], t.compression)
    o = R.TDigest(comp)
    o.absorb_sorted(items, mm[0], mm[1])
    assert t.centroids().tobytes() == o.centroids().tobytes()
    assert t.min() == xa.min() and t.max() == xa.max()
    assert abs(t.size() - wa.sum()) <= 1e-12 * wa.sum()
    return t


@pytest.mark.parametrize("dist", ["lognormal_w", "normal", "uniform", "gamma", "ties", "zeros",
                                  "narrow", "bimodal", "negative", "nonfinite"])
def test_partition_restated_on_cpu(cb, dist):
    rng = np.random.default_rng(17)
    n = 300000
    w = 1.0
    if dist == "lognormal_w":
        x, w = rng.lognormal(size=n), rng.uniform(0.5, 1.5, n)
    elif dist == "normal":
        x = rng.normal(size=n)
    elif dist == "uniform":
        x = rng.uniform(0, 1, n)
    elif dist == "gamma":
        x = rng.gamma(0.1, 0.1, n)
    elif dist == "ties":
        x, w = rng.integers(0, 7, n).astype(float), rng.uniform(0.5, 1.5, n)
    elif dist == "zeros":
        x = rng.choice([0.0, -0.0, 1.0, -1.0, 1e-300, -1e-300], n)
    elif dist == "narrow":
        x = rng.normal(50, 1e-9, n)
    elif dist == "bimodal":
        x = np.concatenate([rng.normal(-1e6, 1e-3, n // 2), rng.normal(1e6, 1e-3, n - n // 2)])
        rng.shuffle(x)
    elif dist == "negative":
        x, w = -rng.lognormal(size=n), rng.uniform(0.5, 1.5, n)
    else:
        x = rng.normal(size=n)
        x[::7] = np.nan
        x[3::11] = np.inf
        x[5::13] = -np.inf
    check_partition(cb, x, w)


@pytest.mark.parametrize("n", [8192, 8193, 10007, 65536, 65537, 100003, 262144 + 255, 1 << 20])
def test_ragged_sizes_and_unaligned_views(cb, n):
    rng = np.random.default_rng(n)
    x = rng.lognormal(size=n + 1)
    w = rng.uniform(0.5, 1.5, n + 1)
    check_partition(cb, x[:n], w[:n], device=True)
    # 8-byte (not 16-byte) aligned device views take the scalar-load variant of the kernel
    dx, dw = cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w)

    class View:
        def __init__(self, d, off, m):
            self.__cuda_array_interface__ = {"shape": (m,), "typestr": "<f8", "version": 3,
                                             "data": (d.ptr + 8 * off, False), "strides": None}
            self._keep = d
    t = cb.TDigest(100, mode="parallel")
    t.update(View(dx, 1, n), View(dw, 1, n))
    u = cb.TDigest(100, mode="parallel")
    u.update(x[1:], w[1:])
    assert t.centroids().tobytes() == u.centroids().tobytes()


@pytest.mark.parametrize("comp", [20, 100, 200, 500, 1000])
def test_compressions(cb, comp):
    rng = np.random.default_rng(comp)
    n = 400000
    check_partition(cb, rng.lognormal(size=n), rng.uniform(0.5, 1.5, n), comp=float(comp))


def test_rank_error_not_worse_than_reference(cb):
    """(b) on config-2 data (lognormal x, U(.5,1.5) w, c=100), 1e7-value subset."""
    rng = np.random.default_rng(1234)
    n = 10**7
    x = rng.lognormal(0, 1, n)
    w = rng.uniform(0.5, 1.5, n)
    rank = weighted_rank_fn(x, w)
    ref = R.TDigest(100)
    ref.update(x, w)
    err_ref = np.abs(rank(ref.quantile(QS9)) - QS9).max()
    t = cb.TDigest(100, mode="parallel")
    t.update(x, w)
    err = np.abs(rank(t.quantile(QS9)) - QS9).max()
    assert err <= max(err_ref, 2.5e-4), (err, err_ref)
    assert err <= 0.012
    xs = np.quantile(x, QS9)
    assert np.abs(t.cdf(xs) - rank(xs)).max() <= max(np.abs(ref.cdf(xs) - rank(xs)).max(), 2.5e-4)
    assert t.min() == x.min() and t.max() == x.max()
    assert abs(t.size() - w.sum()) <= 1e-12 * w.sum()
    # sharded: 8 partial digests merged in rank order (the multi-GPU reduction)
    parts, refs = [], []
    for k in range(8):
        sl = slice(k * n // 8, (k + 1) * n // 8)
        p = cb.TDigest(100, mode="parallel")
        p.update(x[sl], w[sl])
        parts.append(p)
        r = R.TDigest(100)
        r.update(x[sl], w[sl])
        refs.append(r)
    m = cb.TDigest(100)
    m.merge(*parts)
    mr = R.TDigest(100)
    mr.merge(*refs)
    e8 = np.abs(rank(m.quantile(QS9)) - QS9).max()
    e8r = np.abs(rank(mr.quantile(QS9)) - QS9).max()
    assert e8 <= max(2 * e8r, 1e-3), (e8, e8r)


def test_incremental_updates_accumulate(cb):
    rng = np.random.default_rng(6)
    x = rng.normal(size=600000)
    t = cb.TDigest(100, mode="parallel")
    for k in range(3):
        t.update(x[k * 200000:(k + 1) * 200000])
    assert t.size() == 600000 and t.min() == x.min() and t.max() == x.max()
    rank = weighted_rank_fn(x, 1.0)
    assert np.abs(rank(t.quantile(QS9)) - QS9).max() < 1e-3
    a = cb.TDigest(100, mode="auto")     # auto: reference below 8192 values, parallel above
    a.update(x[:1000])
    a.update(x)
    assert a.size() == 601000


def test_full_size_properties_on_device(cb):
    """(c) BASELINE configs[1] at full size on one GPU: 1e9 lognormal + weights, generated on
    the device by the counter-based generator (element i depends on (seed, i) only)."""
    from crick_b200._lib import lib
    n = 10**9
    dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
    lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None)
    t = cb.TDigest(100, mode="parallel")
    t.update(dx, dw)
    # weights are U(0.5, 1.5): total within 5 sigma of n; exact value from a stats pass
    s = cb.SummaryStats()
    s.update(dw)
    assert s.count() == n
    assert abs(t.size() - s.sum()) <= 1e-9 * s.sum()
    sx = cb.SummaryStats()
    sx.update(dx)
    assert t.min() == sx.min() and t.max() == sx.max()
    q = np.linspace(0.0005, 0.9995, 2000)
    est = t.quantile(q)
    assert (np.diff(est) > 0).all()
    # lognormal(0,1): quantile(q) = exp(ndtri(q)); rank error of the estimate in q-space
    from scipy.special import ndtr
    assert np.abs(ndtr(np.log(est)) - q).max() < 5e-4
    # linearity: a second pass doubles every weight and leaves the means where they were
    c1 = t.centroids()
    t2 = cb.TDigest(100, mode="parallel")
    t2.update(dx, dw)
    t2.update(dx, dw)
    assert abs(t2.size() - 2 * t.size()) <= 1e-9 * t.size()
    # (a digest that absorbed two batches has merged centroids: the reference's own merges sit
    # at 2e-4 .. 5e-4 on this data; the caps of its tests are 0.012 / 0.005)
    assert np.abs(ndtr(np.log(t2.quantile(QS9))) - QS9).max() < 2e-3
    assert len(c1) > 90


def test_weights_validated_on_the_device(cb):
    """tdigest.pyx:304-305 for the parallel path: the check runs inside the accumulate kernel;
    a bad weight raises ValueError and leaves the digest as it was."""
    rng = np.random.default_rng(12)
    n = 300000
    x = rng.normal(size=n)
    w = rng.uniform(0.5, 1.5, n)
    t = cb.TDigest(100, mode="parallel")
    t.update(x, w)
    before = t.centroids().tobytes(), t.size()
    for bad in (0.0, -1.0, np.nan, np.inf):
        wb = w.copy()
        wb[n // 3] = bad
        with pytest.raises(ValueError):
            t.update(x, wb)
        with pytest.raises(ValueError):
            t.update(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(wb))
        assert (t.centroids().tobytes(), t.size()) == before
    # tiny positive weights are legal and dropped by the kernel (tdigest_stubs.c:285)
    wb = w.copy()
    wb[::2] = 1e-300
    t2 = cb.TDigest(100, mode="parallel")
    t2.update(x, wb)
    assert abs(t2.size() - wb[1::2].sum()) <= 1e-12 * wb.sum()
    # check_w=False: asynchronous, no validation
    t3 = cb.TDigest(100, mode="parallel")
    t3.update(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w), check_w=False)
    assert abs(t3.size() - w.sum()) <= 1e-12 * w.sum()


def test_deferred_weight_check_needs_no_sync_and_is_sticky(cb):
    """check_w="deferred": the same validation with the verdict kept on the device -- a refused
    update leaves the digest untouched and the ValueError surfaces at the next synchronising
    call; the accepted update is bit-identical with the synchronous form."""
    rng = np.random.default_rng(13)
    n = 200000
    x = rng.lognormal(size=n)
    w = rng.uniform(0.5, 1.5, n)
    dx, dw = cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w)
    t = cb.TDigest(100, mode="parallel")
    t.update(dx, dw, check_w="deferred")
    ref = cb.TDigest(100, mode="parallel")
    ref.update(dx, dw)
    good = t.centroids()
    assert good.tobytes() == ref.centroids().tobytes()
    for bad in (0.0, -2.0, np.nan, np.inf):
        wb = w.copy()
        wb[n // 2] = bad
        t.update(dx, cb.DeviceArray.from_numpy(wb), check_w="deferred")
        t.update(dx, dw, check_w="deferred")          # a later good update does not clear the verdict
        with pytest.raises(ValueError):
            t.centroids()
        t.check_weights()                              # cleared by the raise above
    ref.update(dx, dw); ref.update(dx, dw); ref.update(dx, dw); ref.update(dx, dw)
    assert t.centroids().tobytes() == ref.centroids().tobytes()
    with pytest.raises(ValueError):
        t.update(dx, dw, check_w="later")
    # the scalar weight tdigest_add drops (w <= eps) drops everything (test_tdigest.py:183-193)
    t4 = cb.TDigest()
    t4.update(x, np.finfo("f8").eps)
    t4.update(dx, np.finfo("f8").eps)
    assert t4.size() == 0 and len(t4.centroids()) == 0


def test_sharded_update_over_peer_memory_world1(cb):
    """ShardedTDigest (k_exchange_merge) with one rank: the exchange degenerates to the rank's own
    buffer, and the digest must be the one the plain parallel update of the same data gives in
    size / min / max, with a rank error no worse than the reference's."""
    from crick_b200.distributed import ShardedTDigest
    rng = np.random.default_rng(21)
    n = 600000
    x = rng.lognormal(size=n)
    w = rng.uniform(0.5, 1.5, n)
    dx, dw = cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w)
    sh = ShardedTDigest(100)
    t = sh.update(dx, dw)
    rank = weighted_rank_fn(x, w)
    ref = oracle.reference.TDigest(100) if oracle.reference is not None else R.TDigest(100)
    ref.update(x, w)
    e_ref = np.abs(rank(ref.quantile(QS9)) - QS9).max()
    assert t.min() == x.min() and t.max() == x.max()
    assert abs(t.size() - w.sum()) <= 1e-12 * w.sum()
    assert np.abs(rank(t.quantile(QS9)) - QS9).max() <= max(e_ref, 3e-4)
    c1 = t.centroids().copy()
    # a second step accumulates; unweighted; an empty shard still takes part in the exchange
    sh.update(dx)
    sh.update(cb.DeviceArray((0,)))
    assert abs(sh.digest.size() - (w.sum() + n)) <= 1e-12 * (w.sum() + n)
    # bad weights: the step contributes nothing and the verdict surfaces later
    sh2 = ShardedTDigest(100)
    sh2.update(dx, dw)
    wb = w.copy(); wb[5] = -1.0
    sh2.update(dx, cb.DeviceArray.from_numpy(wb))
    with pytest.raises(ValueError):
        sh2.digest.centroids()
    assert sh2.digest.centroids().tobytes() == c1.tobytes()
    sh.close(); sh2.close()


def _tail_errors(t, x, w, qs):
    tot = w.sum() if np.ndim(w) else float(w) * x.size
    out = []
    for v, q in zip(t.quantile(qs), qs):
        below = w[x < v].sum() if np.ndim(w) else float(w) * np.count_nonzero(x < v)
        upto = w[x <= v].sum() if np.ndim(w) else float(w) * np.count_nonzero(x <= v)
        out.append(abs((below + upto) / (2.0 * tot) - q))
    return np.array(out)


@pytest.mark.parametrize("case", ["lognormal_w_c100_1e8", "lognormal_w_c500", "gamma01_c100"])
def test_tail_rank_error_against_the_reference(cb, case):
    """A t-digest is bought for its tails: weighted rank error at q = 1e-5, 1e-4, 1e-3 and their
    mirrors against the reference's own digest of the SAME data (oracle.reference, the reference's
    C code; one thread, 3-10 s).  Bars, with the reason for each factor:
      * compression 100: within 3x of the reference's error or 2e-5 absolute (measured 0.5-1.9x);
      * compression 500: the main 16 Ki sample resolves nothing below q ~ 1 / 16384 = 6e-5, so the
        far-tail bucket edges come from the 256 Ki-draw tail oversample (k_sort_runs tail groups,
        k_make_edges a'): measured <= 6e-6 at every q here, where the reference (whose first
        centroids weigh ~1e-6 of the stream) reaches 0.4-2.5e-6.  Stated bar: 3x or 1.5e-5 absolute
        (without the oversample q = 0.9999 was at 2.3e-5 and this case failed);
      * gamma(0.1): 4x or 1e-4 (ours is 2-14x BETTER on the lower tail, 1.6-3.2x worse on the upper)."""
    from crick_b200._lib import lib
    Rf = oracle.reference if oracle.reference is not None else R
    qs = np.array([1e-5, 1e-4, 1e-3, 1 - 1e-3, 1 - 1e-4, 1 - 1e-5])
    if case == "gamma01_c100":
        n, comp, factor, floor = 3 * 10**7, 100.0, 4.0, 1e-4
        x = np.random.default_rng(7).gamma(0.1, 1.0, n)
        w = 1.0
        dx, dw = cb.DeviceArray.from_numpy(x), None
    else:
        n = 10**8 if case.endswith("1e8") else 3 * 10**7
        comp = 500.0 if "c500" in case else 100.0
        factor, floor = 3.0, (1.5e-5 if comp == 500.0 else 2e-5)
        dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
        lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None)
        x, w = dx.to_numpy(), dw.to_numpy()
    t = cb.TDigest(comp, mode="parallel")
    if dw is not None:
        t.update(dx, dw)
    else:
        t.update(dx)
    r = Rf.TDigest(comp)
    r.update(x, w if np.ndim(w) else np.ones(n))
    ours, ref = _tail_errors(t, x, w, qs), _tail_errors(r, x, w, qs)
    assert (ours <= np.maximum(factor * ref, floor)).all(), (case, ours.tolist(), ref.tolist())
    # and the 9 standard q's as everywhere else
    assert _tail_errors(t, x, w, QS9).max() <= max(_tail_errors(r, x, w, QS9).max(), 3e-4)


def test_clustered_data_takes_the_binary_search_lookup(cb):
    """Mass in a few narrow, far-apart clusters puts many bucket edges into single lookup cells
    (kmax > 8): the hot kernel then searches the cell's edges instead of counting them.  The path
    must be taken (kmax from the scratch header), restate like every other partition, and not fall
    off a cliff: it is timed here against the ordinary path on the same amount of data."""
    import ctypes
    import time
    from crick_b200._lib import lib
    rng = np.random.default_rng(31)
    n = 1 << 24
    centers = np.array([1e-9, 1.0, 1e3, 1e6, 1e9, 1e12])
    x = (centers[rng.integers(0, len(centers), n)] * (1.0 + 1e-9 * rng.standard_normal(n)))
    w = rng.uniform(0.5, 1.5, n)
    t = check_partition(cb, x, w, device=True)
    raw = (ctypes.c_int * 64)()
    assert lib.crick_tdigest_debug_header(t._h, raw) == 0
    kmax = raw[10]
    assert kmax > 8, "expected the binary-search lookup (kmax = %d)" % kmax
    dx, dw = cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w)
    y = cb.DeviceArray.from_numpy(rng.lognormal(size=n))

    def timed(a):
        d = cb.TDigest(100, mode="parallel")
        d.update(a, dw, check_w=False)
        lib.crick_device_synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            d.update(a, dw, check_w=False)
        lib.crick_device_synchronize()
        return (time.perf_counter() - t0) / 5
    slow, fast = timed(dx), timed(y)
    assert slow < 4.0 * fast, "binary-search lookup %.3f ms vs %.3f ms" % (slow * 1e3, fast * 1e3)


def test_fused_moments_in_the_same_pass(cb):
    """north_star: SummaryStats moments fused into the t-digest pass.  The digest must be the one
    the plain update builds (bit for bit) and the moments must match the stand-alone stats pass
    and NumPy/SciPy at the reference's own tolerance (test_stats.py:20-21)."""
    import scipy.stats
    rng = np.random.default_rng(14)
    n = 1500003
    x = rng.lognormal(size=n)
    x[::5000] = np.nan
    x[7::9000] = np.inf
    w = rng.uniform(0.5, 1.5, n)
    for dev in (False, True):
        t, s = cb.TDigest(100, mode="parallel"), cb.SummaryStats()
        if dev:
            t.update(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w), stats=s)
        else:
            t.update(x, w, stats=s)
        p = cb.TDigest(100, mode="parallel")
        p.update(x, w)
        # (the fused kernel runs 12 warps per SM instead of 14, so the bucket sums are added in
        # another order: same centroids up to rounding, not bit for bit)
        a, b = t.centroids(), p.centroids()
        assert len(a) == len(b)
        np.testing.assert_allclose(a["mean"], b["mean"], rtol=1e-11)
        np.testing.assert_allclose(a["weight"], b["weight"], rtol=1e-11)
        assert abs(t.size() - p.size()) <= 1e-12 * p.size()
        assert t.min() == p.min() and t.max() == p.max()
        ok = x[~np.isnan(x)]
        assert s.count() == ok.size and s.min() == ok.min() and s.max() == np.inf
        fin = np.where(np.isinf(x), 1.0, x)          # moments on a finite copy
        t2, s2 = cb.TDigest(100, mode="parallel"), cb.SummaryStats()
        t2.update(fin, w, stats=s2)
        okf = fin[~np.isnan(fin)]
        alone = cb.SummaryStats()
        alone.update(fin)
        assert s2.count() == alone.count() == okf.size
        assert s2.min() == okf.min() and s2.max() == okf.max()
        np.testing.assert_allclose(s2.mean(), okf.mean(), rtol=1e-12)
        np.testing.assert_allclose(s2.var(), okf.var(), rtol=1e-10)
        np.testing.assert_allclose(s2.skew(), scipy.stats.skew(okf), rtol=0, atol=1e-8)
        np.testing.assert_allclose(s2.kurt(), scipy.stats.kurtosis(okf), rtol=0, atol=1e-7)
        np.testing.assert_allclose(s2.kurt(), alone.kurt(), rtol=0, atol=1e-7)
    # a second fused update merges into both sketches; homogeneous bookkeeping survives
    s3, t3 = cb.SummaryStats(), cb.TDigest(100)
    ones = np.ones(70000)
    t3.update(ones, stats=s3)
    assert np.isnan(s3.skew()) and s3.count() == 70000        # homogeneous -> NaN (stats.pyx)
    t3.update(np.full(70000, 2.0), stats=s3)
    assert s3.count() == 140000 and s3.skew() == 0.0 and t3.size() == 140000


def test_tma_staged_single_table_shape():
    """The alternative kernel shape (one table, tiles staged through shared memory by
    cp.async.bulk + mbarrier) is selected with CRICK_ACC_T=1; it must build the same digest up
    to the summation order.  Runs in a subprocess because the shape is read from the environment."""
    import os, subprocess, sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import crick_b200 as cb
rng = np.random.default_rng(3)
x = rng.lognormal(size=700001); w = rng.uniform(0.5, 1.5, x.size)
for c in (100, 200):
    t = cb.TDigest(c, mode="parallel"); t.update(x, w)
    print(c, t.size().hex(), t.min().hex(), t.max().hex(), len(t.centroids()),
          " ".join(float(v).hex() for v in t.quantile([0.01, 0.5, 0.99])))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env_t in ("2", "1"):
        env = dict(os.environ, CRICK_ACC_T=env_t)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                           timeout=240)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l.split() for l in r.stdout.strip().splitlines()])
    for a, b in zip(*outs):
        assert a[0] == b[0] and a[2:5] == b[2:5]                   # min, max, #centroids
        assert abs(float.fromhex(a[1]) - float.fromhex(b[1])) <= 1e-12 * float.fromhex(a[1])
        qa = np.array([float.fromhex(v) for v in a[5:]])
        qb = np.array([float.fromhex(v) for v in b[5:]])
        np.testing.assert_allclose(qa, qb, rtol=1e-9)


def test_bulk_merge_of_gathered_records(cb):
    """The multi-GPU tail on one GPU: 8 shard digests -> 8 records -> merged (a) with the
    reference's incremental tdigest_merge and (b) in one greedy pass over the sorted union.
    Both must keep size/min/max exactly and sit in the same rank-error class as 8 reference
    digests merged in rank order."""
    rng = np.random.default_rng(1234)
    n = 4 * 10**6
    x = rng.lognormal(0, 1, n)
    w = rng.uniform(0.5, 1.5, n)
    rank = weighted_rank_fn(x, w)
    parts, refs = [], []
    nd = cb.TDigest(100).record_doubles()
    recs = cb.DeviceArray((8, nd))
    for k in range(8):
        sl = slice(k * n // 8, (k + 1) * n // 8)
        p = cb.TDigest(100, mode="parallel")
        p.update(x[sl], w[sl])
        p.export_record(recs.ptr + 8 * nd * k)
        parts.append(p)
        r = R.TDigest(100)
        r.update(x[sl], w[sl])
        refs.append(r)
    from crick_b200._lib import lib
    lib.crick_device_synchronize()
    exact, bulk = cb.TDigest(100), cb.TDigest(100)
    exact.merge_records(recs.ptr, 8)
    bulk.merge_records(recs.ptr, 8, bulk=True)
    via_merge = cb.TDigest(100)
    via_merge.merge(*parts)
    assert exact.centroids().tobytes() == via_merge.centroids().tobytes()
    mr = R.TDigest(100)
    mr.merge(*refs)
    e_ref = np.abs(rank(mr.quantile(QS9)) - QS9).max()
    for t in (exact, bulk):
        assert t.min() == x.min() and t.max() == x.max()
        assert abs(t.size() - w.sum()) <= 1e-12 * w.sum()
        c = t.centroids()
        assert (np.diff(c["mean"]) >= 0).all() and abs(c["weight"].sum() - w.sum()) <= 1e-9 * w.sum()
        err = np.abs(rank(t.quantile(QS9)) - QS9).max()
        assert err <= max(2 * e_ref, 1e-3), (err, e_ref)
    e_bulk = np.abs(rank(bulk.quantile(QS9)) - QS9).max()
    e_exact = np.abs(rank(exact.quantile(QS9)) - QS9).max()
    assert e_bulk <= max(1.5 * e_exact, 5e-4), (e_bulk, e_exact)
    # a second bulk merge accumulates on top of the first
    bulk.merge_records(recs.ptr, 8, bulk=True)
    assert abs(bulk.size() - 2 * w.sum()) <= 1e-12 * w.sum()


@pytest.mark.parametrize("case", ["all_nan", "constant", "two_values", "outlier", "tiny_weights_only",
                                  "int_input", "strided_2d", "sorted", "denormals"])
def test_degenerate_and_awkward_inputs(cb, case):
    rng = np.random.default_rng(99)
    n = 150000
    w = 1.0
    if case == "all_nan":
        x = np.full(n, np.nan)
    elif case == "constant":
        x = np.full(n, 3.25)
    elif case == "two_values":
        x = rng.choice([1.0, 2.0], n, p=[0.999, 0.001])
    elif case == "outlier":
        x = rng.normal(0, 1e-9, n)
        x[n // 2] = 1e300
        x[n // 3] = -1e300
    elif case == "tiny_weights_only":
        x, w = rng.normal(size=n), np.full(n, 1e-300)
    elif case == "int_input":
        x = rng.integers(-1000, 1000, n)
    elif case == "strided_2d":
        x = rng.normal(size=(300, 1000))[:, ::2].T          # non-contiguous: K-order flattening
    elif case == "sorted":
        x = np.sort(rng.lognormal(size=n))
    else:
        x = rng.uniform(0, 1, n) * 1e-310
    t = cb.TDigest(100, mode="parallel")
    t.update(x, w)
    xa, wa = accepted(np.asarray(x, dtype=np.float64).ravel(), w)
    if xa.size == 0:
        assert t.size() == 0 and len(t.centroids()) == 0 and np.isnan(t.min())
        t.update(rng.normal(size=n))                         # still usable afterwards
        assert t.size() == n
        return
    assert t.size() == pytest.approx(wa.sum(), rel=1e-12)
    assert t.min() == xa.min() and t.max() == xa.max()
    c = t.centroids()
    assert (np.diff(c["mean"]) >= 0).all() and (c["weight"] > 0).all()
    rank = weighted_rank_fn(xa, wa)
    qs = np.array([0.01, 0.25, 0.5, 0.75, 0.99])
    est = t.quantile(qs)
    assert (np.diff(est) >= 0).all()
    # rank error at the caps of the reference's own tests (test_tdigest.py:102), ties aside
    if case not in ("constant", "two_values", "int_input"):
        assert np.abs(rank(est) - qs).max() <= 0.012
    else:
        lo = np.searchsorted(np.sort(xa), est, side="left") / xa.size
        hi = np.searchsorted(np.sort(xa), est, side="right") / xa.size
        assert ((lo - 0.012 <= qs) & (qs <= hi + 0.012)).all()


def test_pinned_and_pageable_host_inputs_agree(cb):
    """Host arrays reach the GPU either straight from pinned memory (16 Mi-value chunks) or,
    when pageable, through the library's threaded staging slots (4 Mi-value chunks).  Same
    digest up to the summation order of the chunks; several chunks in both paths."""
    rng = np.random.default_rng(77)
    n = 18 * 10**6 + 12345
    x = rng.lognormal(size=n)
    w = rng.uniform(0.5, 1.5, n)
    px, pw = cb.PinnedArray(n), cb.PinnedArray(n)
    px.array[:] = x
    pw.array[:] = w
    a, b = cb.TDigest(100, mode="parallel"), cb.TDigest(100, mode="parallel")
    a.update(x, w)
    b.update(px.array, pw.array)
    assert a.min() == b.min() == x.min() and a.max() == b.max() == x.max()
    assert abs(a.size() - b.size()) <= 1e-12 * b.size()
    ca, cb_ = a.centroids(), b.centroids()
    assert len(ca) == len(cb_)
    np.testing.assert_allclose(ca["mean"], cb_["mean"], rtol=1e-10)
    np.testing.assert_allclose(ca["weight"], cb_["weight"], rtol=1e-10)
    s = cb.SummaryStats()
    c = cb.TDigest(100, mode="parallel")
    c.update(x, w, stats=s)                       # fused moments across staged chunks
    assert s.count() == n and s.min() == x.min() and s.max() == x.max()
    np.testing.assert_allclose(s.mean(), x.mean(), rtol=1e-12)
    np.testing.assert_allclose(s.var(), x.var(), rtol=1e-9)


def test_pending_buffer_is_flushed_before_a_parallel_update(cb):
    """The library skips the flush launch when it knows the insertion buffer is empty; every
    way of putting items into the buffer (add, small reference-mode update, merge, pickle
    round trip) must still be seen by the next parallel update."""
    rng = np.random.default_rng(77)
    big = rng.normal(size=50000)

    def total_after(prepare):
        t = cb.TDigest(100)
        t.update(big, mode="parallel")          # buffer known empty from here on
        w = prepare(t)
        t.update(big, mode="parallel")
        return t, 2 * big.size + w

    def by_add(t):
        for v in (1.0, 2.0, 3.0):
            t.add(v, 2.0)
        return 6.0

    def by_small_update(t):
        t.update(np.arange(10.0), mode="reference")
        return 10.0

    def by_merge(t):
        o = cb.TDigest(100)
        o.update(np.arange(30.0), mode="reference")
        t.merge(o)
        return 30.0

    for prepare in (by_add, by_small_update, by_merge):
        t, want = total_after(prepare)
        assert t.size() == want, prepare.__name__
        assert t.centroids()["weight"].sum() == want, prepare.__name__
        t2 = pickle.loads(pickle.dumps(t))
        t2.add(5.0)
        t2.update(big, mode="parallel")
        assert t2.centroids()["weight"].sum() == want + 1 + big.size


def test_parallel_update_is_deterministic(cb):
    """Ordered sums + a setup chain launched with programmatic dependencies (griddepcontrol):
    repeating the same updates must reproduce the same bytes every time."""
    rng = np.random.default_rng(123)
    for n in (20000, 700000):
        x = cb.DeviceArray.from_numpy(rng.lognormal(size=n))
        w = cb.DeviceArray.from_numpy(rng.uniform(0.5, 1.5, n))
        seen = set()
        for _ in range(40):
            t = cb.TDigest(100)
            t.update(x, w, mode="parallel")
            t.update(x, mode="parallel")
            seen.add(t.centroids().tobytes())
        assert len(seen) == 1, n


@pytest.mark.gpu
def test_pipelined_updates_equal_plain_updates(cb):
    """TDigest.update(..., tail_stream=) -- sample phase and pass on one stream, the fold into the digest
    on a second one, alternating scratch sets -- leaves the same digest, byte for byte, as the plain
    deferred-check call when it uses the same grid (reserve 0): over back-to-back updates, with a
    query and a scalar add in between, with array and scalar weights; a bad weight is reported by
    check_weights() and skips only its update.  With the default reserve (2 SMs left to the tail
    stream) the bucket sums are cut differently: same quantiles to 1e-6 relative, and the same bytes
    from one run to the next."""
    from crick_b200._lib import lib
    rng = np.random.default_rng(11)
    prev_reserve = lib.crick_pipeline_reserve_sms(0)
    n = 1 << 22
    chunks = [rng.lognormal(0, 1 + 0.2 * i, n) for i in range(5)]
    ws = [rng.uniform(0.5, 2.0, n) for _ in range(5)]
    s1, s2 = lib.crick_stream_create(), lib.crick_stream_create()
    try:
        a, b = cb.TDigest(100, mode="parallel"), cb.TDigest(100, mode="parallel")
        dev = [(cb.DeviceArray.from_numpy(x), cb.DeviceArray.from_numpy(w)) for x, w in zip(chunks, ws)]
        for i, (dx, dw) in enumerate(dev):
            if i == 3:                       # scalar weight, after a host-side add and a query
                for t in (a, b):
                    t.add(0.125, 3.0)
                assert a.quantile(0.5) == b.quantile(0.5)
                a.update(dx, 2.0, mode="parallel", stream=s1)
                b.update(dx, 2.0, stream=s1, tail_stream=s2)
            else:
                a.update(dx, dw, mode="parallel", stream=s1, check_w="deferred")
                b.update(dx, dw, stream=s1, tail_stream=s2)
        a.check_weights(); b.check_weights()
        assert a.centroids().tobytes() == b.centroids().tobytes()
        assert (a.size(), a.min(), a.max()) == (b.size(), b.min(), b.max())
        before = b.centroids().tobytes()
        wbad = ws[0].copy(); wbad[12345] = -1.0
        b.update(dev[0][0], cb.DeviceArray.from_numpy(wbad), stream=s1, tail_stream=s2)
        b.update(dev[1][0], dev[1][1], stream=s1, tail_stream=s2)
        with pytest.raises(ValueError):
            b.check_weights()
        a.update(dev[1][0], dev[1][1], mode="parallel", stream=s1, check_w="deferred")
        assert b.centroids().tobytes() == a.centroids().tobytes() != before
        lib.crick_pipeline_reserve_sms(prev_reserve)
        assert lib.crick_pipeline_reserve_sms(-1) == prev_reserve
        runs = []
        for _ in range(2):
            c = cb.TDigest(100, mode="parallel")
            for dx, dw in dev:
                c.update(dx, dw, stream=s1, tail_stream=s2)
            runs.append(c.centroids().tobytes())
            d = cb.TDigest(100, mode="parallel")
            for dx, dw in dev:
                d.update(dx, dw, mode="parallel", stream=s1)
            qs = np.array([0.001, 0.01, 0.25, 0.5, 0.75, 0.99, 0.999])
            np.testing.assert_allclose(c.quantile(qs), d.quantile(qs), rtol=1e-6)
            assert abs(c.size() - d.size()) <= 1e-12 * d.size()
        assert runs[0] == runs[1]
        # (the overlapped sample phase, crick_pipeline_overlap(1), is not asserted here: it is
        # non-deterministic on updates this small -- tests/tools/pipe_race.py -- and stays an
        # opt-in switch that bench.py does not use)
    finally:
        lib.crick_pipeline_reserve_sms(prev_reserve)
        lib.crick_stream_synchronize(s1); lib.crick_stream_synchronize(s2)
        lib.crick_stream_destroy(s1); lib.crick_stream_destroy(s2)
