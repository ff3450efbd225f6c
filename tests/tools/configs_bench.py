"""Timing of BASELINE.json configs 3, 4 and 5 (parity-test cases, not the bench line) on one
GPU, next to the reference's C code on the host.  Prints one JSON object per config.
Test tooling (it imports oracle/ as checker and CPU baseline), hence under tests/."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crick_b200 as cb
from crick_b200._lib import lib
import oracle

REF = oracle.reference or oracle.restated


def sync():
    lib.crick_device_synchronize()


def timed(fn, reps=3):
    fn(); sync()
    best = 1e30
    for _ in range(reps):
        sync(); t0 = time.perf_counter(); fn(); sync(); best = min(best, time.perf_counter() - t0)
    return best


def config3(ng=10**6, row=1000):
    d = cb.DeviceArray((ng, row))
    lib.crick_synth_fill(d.ptr, None, ng * row, 2, 0, 2, None)   # normal
    sync()
    out = {}
    def run():
        g = cb.TDigestGroup(ng, 100)
        g.update(d)
        g.flush()
        out["g"] = g
    t = timed(run, 2)
    q = out["g"].quantile(cb.DeviceArray.from_numpy(np.array([0.5, 0.99])))
    host = d.to_numpy()[:2000] if ng >= 2000 else d.to_numpy()
    t0 = time.perf_counter()
    for r in range(len(host)):
        o = REF.TDigest(100); o.update(host[r]); o.centroids()
    tc = (time.perf_counter() - t0) / len(host)
    bad = 0
    for r in range(0, len(host), 97):
        o = REF.TDigest(100); o.update(host[r])
        bad += out["g"].centroids(r).tobytes() != o.centroids().tobytes()
    print(json.dumps({"config": 3, "groups": ng, "row": row, "gpu_s": t, "gpu_Gval_s": ng * row / t / 1e9,
                      "bytes_per_digest": row * 8 + 2080, "gpu_GBs": ng * (row * 8 + 2080) / t / 1e9,
                      "cpu_us_per_digest_1core": tc * 1e6, "cpu_Mval_s_1core": row / tc / 1e6,
                      "bit_exact_mismatches": bad}), flush=True)


def config4(nd=65536, per=2000, nq=10**7):
    d = cb.DeviceArray((nd, per))
    lib.crick_synth_fill(d.ptr, None, nd * per, 3, 0, 2, None)
    sync()
    g = cb.TDigestGroup(nd, 200)
    g.update(d); g.flush(); sync()
    ncen = g.meta()[:, 0].sum()
    out = {}
    def merge():
        g2 = cb.TDigestGroup(nd, 200)
        g2.update(d); g2.flush(); sync()
        t0 = time.perf_counter()
        out["m"] = g2.merge_tree(); sync()
        out["t"] = time.perf_counter() - t0
    merge(); merge()
    m = out["m"]
    rng = np.random.default_rng(3)
    q = cb.DeviceArray.from_numpy(rng.uniform(0, 1, nq)); x = cb.DeviceArray.from_numpy(rng.normal(size=nq))
    oq = cb.DeviceArray((nq,))
    tq = timed(lambda: m.quantile(q, out=oq)); tc = timed(lambda: m.cdf(x, out=oq))
    # reference: 2048 digests merged by the same pairwise tree, queries on 1e6
    sub = d.to_numpy()[:2048]
    ds = []
    for r in range(2048):
        o = REF.TDigest(200); o.update(sub[r]); o.flush(); ds.append(o)
    nref = sum(len(o.centroids()) for o in ds)
    t0 = time.perf_counter(); s = 1
    while s < 2048:
        for k in range(0, 2048, 2 * s):
            ds[k].merge(ds[k + s])
        s *= 2
    tm_ref = time.perf_counter() - t0
    qh = rng.uniform(0, 1, 10**6)
    t0 = time.perf_counter(); ds[0].quantile(qh); tq_ref = time.perf_counter() - t0
    print(json.dumps({"config": 4, "digests": nd, "centroids_in": int(ncen), "merge_tree_s": out["t"],
                      "merge_Mcentroid_s": ncen / out["t"] / 1e6, "merged_centroids": len(m.centroids()),
                      "quantile_1e7_s": tq, "quantile_Gq_s": nq / tq / 1e9, "quantile_GBs": 16 * nq / tq / 1e9,
                      "cdf_1e7_s": tc, "cdf_Gq_s": nq / tc / 1e9,
                      "cpu_merge_Mcentroid_s_1core": nref / tm_ref / 1e6, "cpu_quantile_Mq_s_1core": 1e6 / tq_ref / 1e6}), flush=True)


def config5(n=10**8):
    rng = np.random.default_rng(5)
    x = rng.zipf(1.1, n).astype(np.int64)
    dx = cb.DeviceArray.from_numpy(x)
    s = cb.SummaryStats()
    ts = timed(lambda: s.update(dx))
    m = 2 * 10**6
    sp = cb.SpaceSaving(1000, "i8")
    t0 = time.perf_counter(); sp.update(cb.DeviceArray.from_numpy(x[:m])); sp.size(); sync(); tp = time.perf_counter() - t0
    rp = cb.SpaceSaving(1000, "i8")          # in-order replay: the reference's bits, list order included
    t0 = time.perf_counter(); rp.update(cb.DeviceArray.from_numpy(x[:m]), mode=1); rp.size(); sync(); trp = time.perf_counter() - t0
    big = cb.SpaceSaving(1000, "i8")
    t0 = time.perf_counter(); big.update(dx); big.size(); sync(); tbig = time.perf_counter() - t0
    vals, counts = np.unique(x, return_counts=True)
    top_true = set(vals[np.argsort(-counts, kind="stable")[:100]].tolist())
    top_ok = top_true <= set(big.counters()["item"][:300].tolist())
    o = REF.SummaryStats(); t0 = time.perf_counter(); o.update(x[:10**7].astype(np.float64)); tso = time.perf_counter() - t0
    so = REF.SpaceSaving(1000); t0 = time.perf_counter(); so.update(x[:m]); tpo = time.perf_counter() - t0
    print(json.dumps({"config": 5, "stats_Gval_s": n / ts / 1e9, "stats_GBs": 8 * n / ts / 1e9,
                      "spsv_Mval_s_2e6": m / tp / 1e6, "spsv_Mval_s_1e8": n / tbig / 1e6, "spsv_top100_found": bool(top_ok), "cpu_stats_Mval_s_1core": 1e7 / tso / 1e6,
                      "cpu_spsv_Mval_s_1core": m / tpo / 1e6,
                      "spsv_replay_Mval_s_2e6": m / trp / 1e6,
                      "spsv_replay_2e6_equals_reference": rp.counters().tobytes() == so.counters().tobytes()}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5"]
    if "3" in which: config3(int(os.environ.get("CFG3_GROUPS", 10**6)))
    if "4" in which: config4()
    if "5" in which: config5()
