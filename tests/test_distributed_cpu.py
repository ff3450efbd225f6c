"""CPU, world_size 2 over gloo: the host-side plumbing of the multi-GPU reduction
-- shard bounds, the fixed-size record layout, and the rank-order all-gather.  The merge itself
is a device kernel (tests/test_gpu_*.py); here the oracle plays its part so that the gathered
records can be checked end to end against "G reference digests merged in rank order"."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import oracle
from crick_b200.distributed import (gather_records, record_from_state, shard_bounds,
                                    state_from_record)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 200000
rng = np.random.default_rng(1234)
x, w = rng.lognormal(size=n), rng.uniform(0.5, 1.5, n)
lo, hi = shard_bounds(n, rank, world)
d = oracle.restated.TDigest(100)
d.update(x[lo:hi], w[lo:hi])
c, tot, mn, mx = d.__getstate__()
rec = torch.from_numpy(record_from_state(c, tot, mn, mx, 200))
g = gather_records(rec).numpy()
ok = g.shape == (world, 404) and np.array_equal(g[rank], rec.numpy())
# every rank rebuilds all digests from the records and merges them in rank order
parts = []
for r in range(world):
    cen, t_, a_, b_ = state_from_record(g[r])
    p = oracle.restated.TDigest(100)
    p.__setstate__((cen.ravel().view(oracle.CENTROID_DTYPE), t_, a_, b_))
    parts.append(p)
m = oracle.restated.TDigest(100)
m.merge(*parts)
# the expected result, computed locally from the full stream
exp_parts = []
for r in range(world):
    a, b = shard_bounds(n, r, world)
    q = oracle.restated.TDigest(100); q.update(x[a:b], w[a:b]); exp_parts.append(q)
e = oracle.restated.TDigest(100)
e.merge(*exp_parts)
ok = ok and m.centroids().tobytes() == e.centroids().tobytes() and m.size() == e.size()
# SpaceSaving: fixed-size int64 records, gathered and merged in rank order
from crick_b200.distributed import spsv_counters_from_record, spsv_record_from_counters
items = (rng.zipf(1.2, n) % 5000).astype(np.int64)
sp = oracle.restated.SpaceSaving(100)
sp.update(items[lo:hi])
srec = torch.from_numpy(spsv_record_from_counters(sp.counters(), 100))
sg = gather_records(srec).numpy()
ok = ok and sg.shape == (world, 302) and np.array_equal(sg[rank], srec.numpy())
sm_ = oracle.restated.SpaceSaving(100)
for r in range(world):
    c3 = spsv_counters_from_record(sg[r])
    p = oracle.restated.SpaceSaving(100)
    st_ = np.zeros(len(c3), dtype=oracle.COUNTER_DTYPE)
    st_["item"], st_["count"], st_["error"] = c3[:, 0], c3[:, 1], c3[:, 2]
    p.set_state(st_)
    sm_.merge(p)
se = oracle.restated.SpaceSaving(100)
for r in range(world):
    a, b = shard_bounds(n, r, world)
    q = oracle.restated.SpaceSaving(100); q.update(items[a:b]); se.merge(q)
ok = ok and sm_.counters().tobytes() == se.counters().tobytes()
res = torch.tensor([1 if ok else 0])
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": int(res.item()), "world": world, "n_centroids": len(m.centroids())}))
dist.destroy_process_group()
'''


def test_record_roundtrip_and_shard_bounds():
    from crick_b200.distributed import record_from_state, shard_bounds, state_from_record
    n = 10**9 + 7
    cuts = [shard_bounds(n, r, 8) for r in range(8)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    assert all(cuts[r][1] == cuts[r + 1][0] for r in range(7))
    assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    c = np.arange(10.0).reshape(5, 2)
    rec = record_from_state(c, 25.0, -1.0, 9.0, 200)
    assert rec.shape == (404,) and rec[0] == 5
    c2, tot, mn, mx = state_from_record(rec)
    assert np.array_equal(c, c2) and (tot, mn, mx) == (25.0, -1.0, 9.0)


def test_gloo_world2_gather_and_rank_order_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT],
        capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res == {"ok": 1, "world": 2, "n_centroids": res["n_centroids"]} and res["n_centroids"] > 90
