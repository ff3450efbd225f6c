"""Multi-GPU reduction through the C ABI's own NCCL communicator, WITHOUT PyTorch: `world` processes
(multiprocessing spawn, one GPU each), the unique id travels through a pipe.  Every rank builds a
TDigest / SpaceSaving / SummaryStats over its shard, NcclComm reduces them; rank 0 rebuilds every
shard's sketch locally and checks the reduced ones against `merge(*shards)` byte for byte.
usage: comm_check.py [world] [n_per_rank]"""
import json, sys, time
import multiprocessing as mp
import numpy as np

sys.path.insert(0, ".")


def shard(rank, n):
    rng = np.random.default_rng(100 + rank)
    return rng.normal(rank * 0.1, 1.0, n), rng.zipf(1.3, n).astype(np.int64)


def build(cb, x, items):
    t = cb.TDigest(100.0)
    t.update(x)
    s = cb.SpaceSaving(64, "i8")
    s.update(items)
    m = cb.SummaryStats()
    m.update(x)
    return t, s, m


def worker(rank, world, n, conn, q):
    assert "torch" not in sys.modules
    import crick_b200 as cb
    from crick_b200._lib import lib
    from crick_b200.distributed import NcclComm
    lib.crick_set_device(rank)
    if rank == 0:
        uid = NcclComm.unique_id()
        for c in conn:
            c.send(uid)
    else:
        uid = conn.recv()
    comm = NcclComm(uid, rank, world)
    x, items = shard(rank, n)
    t, s, m = build(cb, x, items)
    res = {}
    for exact in (True, False):
        rt = comm.reduce_tdigest(t, exact=exact)
        rs = comm.reduce_space_saving(s, exact=exact)
        rm = comm.reduce_stats(m)
        comm.synchronize()
        res[exact] = (rt.centroids().tobytes(), rt.size(), np.asarray(rs.counters()).tobytes(), rm.count(), rm.mean(), rm.var())
    # latency of one reduction of all three sketches
    out_t, out_s, out_m = cb.TDigest(100.0), cb.SpaceSaving(64, "i8"), cb.SummaryStats()
    for it in range(25):
        if it == 5:
            comm.synchronize(); t0 = time.perf_counter()
        comm.reduce_tdigest(t, out=out_t)
        comm.reduce_space_saving(s, out=out_s)
        comm.reduce_stats(m, out=out_m)
    comm.synchronize()
    us = (time.perf_counter() - t0) / 20 * 1e6
    if rank == 0:
        parts = [build(cb, *shard(r, n)) for r in range(world)]
        et, es, em = cb.TDigest(100.0), cb.SpaceSaving(64, "i8"), cb.SummaryStats()
        et.merge(*[p[0] for p in parts]); es.merge(*[p[1] for p in parts]); em.merge(*[p[2] for p in parts])
        allx = np.concatenate([shard(r, n)[0] for r in range(world)])
        ok_exact = (res[True][0] == et.centroids().tobytes() and res[True][2] == np.asarray(es.counters()).tobytes()
                    and res[True][3] == em.count() and res[True][4] == em.mean() and res[True][5] == em.var())
        q.put({"world": world, "n_per_rank": n, "torch_imported": "torch" in sys.modules,
               "exact_equals_merge_of_shards": bool(ok_exact),
               "bulk_total_weight": float(res[False][1]), "expected_total": float(len(allx)),
               "bulk_topk_equal": res[False][2] == np.asarray(es.counters()).tobytes(),
               "us_per_reduction_of_three_sketches": us})
    comm.close()


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes = [ctx.Pipe() for _ in range(world - 1)]
    procs = [ctx.Process(target=worker, args=(0, world, n, [p[0] for p in pipes], q))]
    procs += [ctx.Process(target=worker, args=(r, world, n, pipes[r - 1][1], q)) for r in range(1, world)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(60)
    print(json.dumps(out))
    sys.exit(0 if out["exact_equals_merge_of_shards"] and not out["torch_imported"] else 1)
