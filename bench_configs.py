"""BASELINE.json configs 3, 4 and 5 at full size, for bench.py's `other_configs` key.

Each entry: value (+ unit), ms, roofline_frac (algorithmic bytes of SURVEY.md 8d / time / measured
HBM peak), parity_mismatches (the CUDA path against the oracle on a bounded sample of the SAME
data -- the oracle is used as the checker and, at N = 1, as the timed CPU baseline; it is never on
the measured path) and cpu_baseline.  With N > 1 the shapes that shard do (config 3 by groups,
config 5 by contiguous element ranges + all-gather merge of the summaries); config 4's merge tree
is one digest set on one GPU and only runs at N = 1.
"""
import time

import numpy as np

QS9 = np.array([0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.99, 0.999])


def _shard(n, rank, world):
    return (rank * n) // world, ((rank + 1) * n) // world


def _timed(B, fn, reps=3):
    """min over `reps` of the wall time of fn() between device synchronisations, max over ranks."""
    lib = B.lib
    fn()
    best = 1e30
    for _ in range(reps):
        B.barrier()
        t0 = time.perf_counter()
        fn()
        lib.crick_device_synchronize()
        best = min(best, time.perf_counter() - t0)
    return B.max_over_ranks(best)


def _ref():
    import oracle
    return oracle.reference if oracle.reference is not None else oracle.restated


def _rows_to_host(B, dev, first_row, nrows, row_len):
    out = np.empty((nrows, row_len), dtype=np.float64)
    B.lib.crick_memcpy_d2h(out.ctypes.data, dev.ptr + first_row * row_len * 8, out.nbytes, None)
    return out


# ---- config 3: 1e6 independent digests x 1000 values, reference-compatible (bit-exact) ----------
def config3(B, cpu, ng_total=10**6, row=1000):
    cb, lib = B.cb, B.lib
    glo, ghi = _shard(ng_total, B.rank, B.world)
    ng = ghi - glo
    d = cb.DeviceArray((ng, row))
    lib.crick_synth_fill(d.ptr, None, ng * row, 2, glo * row, 2, None)      # N(0, 1)
    lib.crick_device_synchronize()
    keep = {}

    def run():
        g = cb.TDigestGroup(ng, 100)
        g.update(d)
        g.flush()
        keep["g"] = g
    import ctypes
    nfb = ctypes.c_int64(0)
    lib.crick_group_exact_fallbacks(ctypes.byref(nfb), 1)
    run()
    lib.crick_group_exact_fallbacks(ctypes.byref(nfb), 1)
    t = _timed(B, run, 4)      # (min of 4: each rep also allocates and frees the 4 GB of digest state)
    g = keep["g"]
    bytes_per_digest = row * 8 + 2080                      # input + ~130 centroids out
    out = {"workload": "grouped: %d digests x %d values, compression 100, reference-compatible mode "
                       "(sharded by groups over %d GPU(s), no exchange)" % (ng_total, row, B.world),
           "value": ng_total * row / t / 1e9, "unit": "Gvalues/s", "ms": t * 1e3,
           "roofline_frac": ng * bytes_per_digest / t / 1e9 / B.peak,
           "algorithmic_bytes_per_digest": bytes_per_digest,
           "exact_kscale_fallbacks_per_pass": int(nfb.value),
           "decision_margin": float(lib.crick_group_margin(float("nan")))}
    if B.rank == 0:
        R = _ref()
        rows = [0, 1, 7, ng // 3, ng // 2, ng - 1] + list(range(100, min(ng, 2100), 97))
        bad = 0
        for r in rows:
            host = _rows_to_host(B, d, r, 1, row)[0]
            o = R.TDigest(100)
            o.update(host)
            bad += g.centroids(r).tobytes() != o.centroids().tobytes()
        out["parity_mismatches"] = int(bad)
        out["parity_sample"] = "%d digests, centroids byte for byte vs oracle.%s" % (len(rows), R.kind)
        if cpu:
            host = _rows_to_host(B, d, 0, 2000, row)
            t0 = time.perf_counter()
            for r in range(len(host)):
                o = R.TDigest(100); o.update(host[r]); o.centroids()
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": len(host) * row / dt / 1e9, "unit": "Gvalues/s", "cores": 1,
                                   "kind": R.kind, "sample": "%d digests of the same data, one core (%.2f s)" % (len(host), dt)}
    del keep["g"], g, d
    return out


# ---- config 4: merge 65536 digests (compression 200), then 1e7 quantile / cdf queries -----------
def config4(B, cpu, nd=65536, per=2000, nq=10**7):
    cb, lib = B.cb, B.lib
    d = cb.DeviceArray((nd, per))
    lib.crick_synth_fill(d.ptr, None, nd * per, 3, 0, 2, None)
    lib.crick_device_synchronize()
    keep = {}

    def build():
        g = cb.TDigestGroup(nd, 200)
        g.update(d)
        g.flush()
        lib.crick_device_synchronize()
        return g
    ncen = float(build().meta()[:, 0].sum())
    out = {"workload": "merge-heavy: %d partial digests (compression 200, %d values each, %.1f M centroids) "
                       "merged into one, then %d quantile and %d cdf queries" % (nd, per, ncen / 1e6, nq, nq)}
    for mode in ("tree", "fast"):
        best = 1e30
        for _ in range(3):
            g = build()
            t0 = time.perf_counter()
            keep[mode] = g.merge_tree() if mode == "tree" else g.merge_all()
            lib.crick_device_synchronize()
            best = min(best, time.perf_counter() - t0)
            del g
        out["merge_" + mode] = {"ms": best * 1e3, "value": ncen / best / 1e9, "unit": "Gcentroids/s",
                                "roofline_frac": ncen * 16 / best / 1e9 / B.peak,
                                "merged_centroids": int(len(keep[mode].centroids()))}
    out["merge_tree"]["what"] = "pairwise level-by-level tree, exactly nested tdigest_merge calls (bit-exact with the reference)"
    out["merge_fast"]["what"] = ("TDigestGroup.merge_all: all centroids as (mean, weight) points through ONE parallel "
                                 "k-bucket update (flush, gather, sample/edges, one streaming pass, one greedy compression)")
    mt = keep["tree"]
    rng = np.random.default_rng(3)
    q = cb.DeviceArray.from_numpy(rng.uniform(0, 1, nq))
    x = cb.DeviceArray.from_numpy(rng.normal(size=nq))
    oq = cb.DeviceArray((nq,))
    tq = _timed(B, lambda: mt.quantile(q, out=oq))
    tc = _timed(B, lambda: mt.cdf(x, out=oq))
    out["quantile"] = {"ms": tq * 1e3, "value": nq / tq / 1e9, "unit": "Gqueries/s",
                       "roofline_frac": 16 * nq / tq / 1e9 / B.peak}
    out["cdf"] = {"ms": tc * 1e3, "value": nq / tc / 1e9, "unit": "Gqueries/s",
                  "roofline_frac": 16 * nq / tc / 1e9 / B.peak}
    out["value"], out["unit"] = out["merge_fast"]["value"], "Gcentroids/s"
    out["ms"] = out["merge_fast"]["ms"]
    out["roofline_frac"] = out["merge_fast"]["roofline_frac"]
    # parity: a 512-digest tree against the oracle's nested merges, bit for bit; queries on it;
    # the fast merge against the exact one by rank error
    R = _ref()
    nsub = 512
    sub = _rows_to_host(B, d, 0, nsub, per)
    gs = cb.TDigestGroup(nsub, 200)
    gs.update(sub)
    gs.flush()
    ms_ = gs.merge_tree()
    ds = []
    for r in range(nsub):
        o = R.TDigest(200); o.update(sub[r]); o.flush(); ds.append(o)
    t0 = time.perf_counter()
    nref = sum(len(o.centroids()) for o in ds)
    s = 1
    while s < nsub:
        for k in range(0, nsub, 2 * s):
            ds[k].merge(ds[k + s])
        s *= 2
    t_merge = time.perf_counter() - t0
    bad = int(ms_.centroids().tobytes() != ds[0].centroids().tobytes())
    qh = rng.uniform(0, 1, 10**5)
    xh = rng.normal(size=10**5)
    bad += int(ms_.quantile(qh).tobytes() != ds[0].quantile(qh).tobytes())
    bad += int(ms_.cdf(xh).tobytes() != ds[0].cdf(xh).tobytes())
    out["parity_mismatches"] = bad
    out["parity_sample"] = ("%d-digest tree vs nested oracle.%s merges (centroids, 1e5 quantiles, 1e5 cdf values: "
                            "byte for byte)" % (nsub, R.kind))
    # the fast merge is judged on rank error against the exact tree's digest (same class required)
    allv = np.sort(d.to_numpy().ravel()[:: 16])
    ex, fa = keep["tree"].quantile(QS9), keep["fast"].quantile(QS9)
    rk = lambda v: np.searchsorted(allv, v) / len(allv)
    out["merge_fast"]["rank_error_max"] = float(np.abs(rk(fa) - QS9).max())
    out["merge_tree"]["rank_error_max"] = float(np.abs(rk(ex) - QS9).max())
    if cpu:
        t0 = time.perf_counter(); ds[0].quantile(rng.uniform(0, 1, 10**6)); tqr = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nref / t_merge / 1e9, "unit": "Gcentroids/s", "cores": 1, "kind": R.kind,
                               "sample": "%d digests merged by the same tree on one core (%.3f s); quantile "
                                         "%.1f Mqueries/s on 1e6 queries" % (nsub, t_merge, 1.0 / tqr)}
    return out


# ---- config 5: SummaryStats moments + SpaceSaving(1000) over 1e9 Zipf(1.1) int64, ONE pass ------
def config5(B, cpu, n_total=10**9, capacity=1000):
    cb, lib = B.cb, B.lib
    lo, hi = _shard(n_total, B.rank, B.world)
    m = hi - lo
    dx = cb.DeviceArray((m,), "i8")
    lib.crick_synth_fill(dx.ptr, None, m, 5, lo, 4, None)                  # Zipf(1.1)-like int64
    lib.crick_device_synchronize()
    keep = {}
    reducer = None
    if B.world > 1:
        from crick_b200.distributed import SketchAllReduce
        reducer = SketchAllReduce(capacity, "i8", exact=False)

    def run():
        sp, st = cb.SpaceSaving(capacity, "i8"), cb.SummaryStats()
        sp.update(dx, stats=st, stream=B.sptr)                              # crick_spsv_update_stats
        keep["local"] = (sp, st)
        if reducer is not None:
            # one all-gather of (summary record | moments record), both merges on the device
            with B.torch.cuda.stream(B.stream):
                sp, st = reducer(sp, st)
        keep["sp"], keep["st"] = sp, st

    def run_two_passes():
        sp, st = cb.SpaceSaving(capacity, "i8"), cb.SummaryStats()
        sp.update(dx)
        st.update(dx)
        keep["sp2"], keep["st2"] = sp, st
    t = _timed(B, run, 3)
    t2 = _timed(B, run_two_passes, 3)
    sp, st = keep["sp"], keep["st"]
    c = sp.counters()
    out = {"workload": "SummaryStats moments + SpaceSaving(%d) over ONE stream of %d Zipf(1.1) int64 items in one "
                       "fused pass (sharded over %d GPU(s) by contiguous ranges%s)"
                       % (capacity, n_total, B.world, ", (summary | moments) records all-gathered once and merged on the device (k-way merge)" if B.world > 1 else ""),
           "value": n_total / t / 1e9, "unit": "Gvalues/s", "ms": t * 1e3,
           "roofline_frac": 8 * m / t / 1e9 / B.peak,
           "two_passes_ms": t2 * 1e3, "fused_over_two_passes": t / t2,
           "stats_count": int(st.count()), "counters": int(len(c)), "max_error": int(c["error"].max()),
           "min_count": int(c["count"].min())}
    if B.rank == 0:
        bad = int(st.count() != n_total) + int((np.diff(c["count"]) > 0).any())
        # fused == separate on this rank's shard (byte for byte for the summary; moments 1e-11)
        if B.world == 1:
            bad += int(c.tobytes() != keep["sp2"].counters().tobytes())
            a, b = st.__getstate__(), keep["st2"].__getstate__()
            bad += int(a[0] != b[0] or a[2] != b[2] or a[3] != b[3])
            bad += int(any(abs(u - v) > 1e-11 * abs(v) for u, v in zip(a[1:7], b[1:7]) if v != 0))
        # exact counts on a prefix the host can count: every heavy hitter of the first 1e7 items
        npre = min(m, 10**7)
        pre = np.empty(npre, dtype=np.int64)
        lib.crick_memcpy_d2h(pre.ctypes.data, dx.ptr, pre.nbytes, None)
        s2 = cb.SpaceSaving(capacity, "i8")
        s2.update(pre)
        vals, counts = np.unique(pre, return_counts=True)
        true = dict(zip(vals.tolist(), counts.tolist()))
        got = {it: (cn, er) for it, cn, er in s2.counters().tolist()}
        top = vals[np.argsort(-counts, kind="stable")[:200]]
        bad += sum(1 for it in top.tolist() if it not in got or got[it][0] - got[it][1] != true[it])
        # exact regime (#distinct <= capacity): the reference's bits, list order included
        R = _ref()
        small = (pre[: 2 * 10**6] % 700) * 7919 - 3
        s3, o3 = cb.SpaceSaving(capacity, "i8"), R.SpaceSaving(capacity)
        s3.update(small)
        o3.update(small)
        bad += int(s3.counters().tobytes() != o3.counters().tobytes())
        out["parity_mismatches"] = int(bad)
        out["parity_sample"] = ("stats count; fused == two passes; top-200 of a 1e7-item prefix exact vs numpy; "
                                "2e6 items with 700 distinct values byte for byte vs oracle.%s" % R.kind)
        if cpu:
            xs = pre.astype(np.float64)
            o = R.SummaryStats(); t0 = time.perf_counter(); o.update(xs); ts = time.perf_counter() - t0
            so = R.SpaceSaving(capacity); t0 = time.perf_counter(); so.update(pre[: 2 * 10**6]); tp = time.perf_counter() - t0
            per_item = ts / len(xs) + tp / (2 * 10**6)
            out["cpu_baseline"] = {"value": 1.0 / per_item / 1e9, "unit": "Gvalues/s", "cores": 1, "kind": R.kind,
                                   "sample": "SummaryStats on 1e7 items (%.0f Mvalues/s) + SpaceSaving on 2e6 items "
                                             "(%.2f Mvalues/s) of the same stream, one core" % (len(xs) / ts / 1e6, 2.0 / tp)}
    keep.clear()
    del dx
    return out


def run_other_configs(B, cpu):
    out = {}
    for name, fn in (("3", config3), ("4", config4), ("5", config5)):
        if name == "4" and B.world > 1:
            out[name] = {"skipped": "one digest set on one GPU: measured at N = 1 only"}
            continue
        try:
            out[name] = fn(B, cpu)
        except Exception as exc:   # a failing side config must not take the headline line with it
            out[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        B.barrier()
    return out
