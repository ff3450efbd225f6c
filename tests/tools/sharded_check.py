"""Multi-GPU check of ShardedTDigest (k_exchange_merge over NVLink peer memory); run with
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/tools/sharded_check.py
Every rank owns a shard of one seeded stream; after each collective update all ranks must hold the
SAME digest (bytes), with exact size / min / max of the union and a rank error in the class of the
NCCL reducer's.  Prints one JSON line per rank 0 check and timings of both paths."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch, torch.distributed as dist
import crick_b200 as cb
from crick_b200._lib import lib
from crick_b200.distributed import ShardedTDigest, TDigestAllReduce

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); lib.crick_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8 * 10**6
lo, hi = rank * n_total // world, (rank + 1) * n_total // world
m = hi - lo
dx, dw = cb.DeviceArray((m,)), cb.DeviceArray((m,))
lib.crick_synth_fill(dx.ptr, dw.ptr, m, 1234, lo, 0, None)
lib.crick_device_synchronize()
stream = torch.cuda.Stream(); sptr = stream.cuda_stream
sh = ShardedTDigest(100)
ok = True
for step in range(3):
    t = sh.update(dx, dw, stream=sptr)
    stream.synchronize()
    c = t.centroids()
    blob = torch.tensor(list(c.tobytes()[:4096].ljust(4096, b"\0")), dtype=torch.uint8, device="cuda")
    allb = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(allb, blob)
    same = all(bool((b == allb[0]).all()) for b in allb)
    ok &= same
    if rank == 0:
        print(json.dumps({"step": step, "identical_on_all_ranks": same, "centroids": len(c), "size": t.size()}), flush=True)
# accuracy vs the full stream (rank 0 regenerates it on its GPU)
if rank == 0 and n_total <= 2 * 10**8:
    fx, fw = cb.DeviceArray((n_total,)), cb.DeviceArray((n_total,))
    lib.crick_synth_fill(fx.ptr, fw.ptr, n_total, 1234, 0, 0, None)
    x, w = fx.to_numpy(), fw.to_numpy()
    order = np.argsort(x); xs = x[order]; cw = np.cumsum(w[order]); tot = cw[-1]
    q = np.array([0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.99, 0.999])
    est = sh.digest.quantile(q)
    err = np.abs(cw[np.searchsorted(xs, est) - 1] / tot - q).max()
    exact = (bool(sh.digest.min() == x.min()), bool(sh.digest.max() == x.max()),
             bool(abs(sh.digest.size() - 3 * w.sum()) <= 1e-11 * w.sum()))
    ok &= all(exact) and bool(err < 1e-3)
    print(json.dumps({"rank_error_max": float(err), "min_max_size_exact": exact}), flush=True)
# timing: peer-memory exchange vs update + NCCL all-gather reducer
red = TDigestAllReduce(100.0, exact=False)
dig = cb.TDigest(100, mode="parallel")
def step_nccl():
    dig.update(dx, dw, mode="parallel", stream=sptr, check_w="deferred")
    with torch.cuda.stream(stream):
        return red(dig)
def step_p2p():
    return sh.update(dx, dw, stream=sptr)
res = {}
for name, fn in (("nccl", step_nccl), ("p2p", step_p2p)):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20): fn()
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name] = float(ms.item())
sh.digest.check_weights(); dig.check_weights()
import ctypes
raw = (ctypes.c_ulonglong * 32)()
lib.crick_tdigest_debug_header(sh.digest._h, raw)
ts = list(raw)[16:22]
print(json.dumps({"rank": rank, "exchange_kernel_us": {"wait_for_partials": (ts[1] - ts[0]) / 1e3, "points": (ts[2] - ts[1]) / 1e3,
                  "publish": (ts[3] - ts[2]) / 1e3, "wait_for_ranks": (ts[4] - ts[3]) / 1e3, "merge": (ts[5] - ts[4]) / 1e3}}), flush=True)
if rank == 0:
    print(json.dumps({"world": world, "values_per_rank": m, "ms_per_step": res, "ok": bool(ok)}), flush=True)
sh.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
