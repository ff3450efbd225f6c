"""One multi-GPU step on ONE GPU (a 1-rank NCCL group): update of an n-value shard -> export ->
all-gather -> merge, as bench.py runs it at N > 1.  Prints ms per step and the hot kernel's share:
the difference is the fixed cost that limits strong scaling.  python tests/tools/step_probe.py [n] [records]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
import torch, torch.distributed as dist
import crick_b200 as cb
from crick_b200._lib import lib
from crick_b200.distributed import TDigestAllReduce

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125000000
torch.cuda.set_device(0); lib.crick_set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None)
lib.crick_device_synchronize()
stream = torch.cuda.Stream(); sptr = stream.cuda_stream
tail = torch.cuda.Stream(); tptr = tail.cuda_stream
PIPE = os.environ.get("PIPE", "1") != "0"
if "OVERLAP" in os.environ: lib.crick_pipeline_overlap(int(os.environ["OVERLAP"]))
if "RESERVE" in os.environ: lib.crick_pipeline_reserve_sms(int(os.environ["RESERVE"]))
red = TDigestAllReduce(100.0, exact=False)
t = cb.TDigest(100, mode="parallel")
def step():
    if PIPE:
        t.update(dx, dw, stream=sptr, tail_stream=tptr)
    else:
        t.update(dx, dw, mode="parallel", stream=sptr, check_w="deferred")
    with torch.cuda.stream(tail if PIPE else stream):
        return red(t)
for _ in range(5): step()
torch.cuda.synchronize()
lib.crick_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record(stream)
for _ in range(K): m = step()
stream.wait_stream(tail)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
lib.crick_profile_enable(0)
kms, kn = ctypes.c_double(0), ctypes.c_int64(0)
lib.crick_profile_collect(ctypes.addressof(kms), ctypes.addressof(kn))
t.check_weights()
print(json.dumps({"pipelined": PIPE, "overlap": lib.crick_pipeline_overlap(-1), "reserve": lib.crick_pipeline_reserve_sms(-1), "n": n, "ms_per_step": ms, "kernel_ms": kms.value / kn.value, "fixed_us": 1e3 * (ms - kms.value / kn.value),
                  "centroids": len(m.centroids()),
                  "centroids_md5": __import__("hashlib").md5(m.centroids().tobytes()).hexdigest()[:12]}), flush=True)
import numpy as np
hdr = np.zeros(32, dtype=np.uint64)
if lib.crick_tdigest_debug_header(t._h, hdr.ctypes.data) == 0:
    st = hdr[16:].astype(np.int64)
    names = {6: "sort0_start", 7: "sort0_end", 8: "tail0_loaded", 9: "tail0_end", 10: "edges_start", 11: "edges_tails",
             12: "edges_end", 13: "hot0_start", 14: "hot0_end", 0: "fin_start", 15: "fin_reduced", 2: "fin_points", 5: "fin_end"}
    t0 = int(st[6])
    print(json.dumps({"timeline_us": {names[k]: round((int(st[k]) - t0) / 1e3, 1) for k in (6, 8, 9, 7, 10, 11, 12, 13, 14, 0, 15, 2, 5)}}))
dist.destroy_process_group()
