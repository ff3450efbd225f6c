"""The sharded-update step (ShardedTDigest / k_exchange_merge) on ONE GPU with a 1-rank
communicator: ms per step and the hot kernel's share.  python tests/tools/step_probe_p2p.py [n]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import crick_b200 as cb
from crick_b200._lib import lib
from crick_b200.distributed import ShardedTDigest

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125000000
torch.cuda.set_device(0); lib.crick_set_device(0)
dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, 0, None)
lib.crick_device_synchronize()
stream = torch.cuda.Stream(); sptr = stream.cuda_stream
sh = ShardedTDigest(100)
for _ in range(5): sh.update(dx, dw, stream=sptr)
torch.cuda.synchronize()
lib.crick_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record(stream)
for _ in range(K): sh.update(dx, dw, stream=sptr)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
lib.crick_profile_enable(0)
kms, kn = ctypes.c_double(0), ctypes.c_int64(0)
lib.crick_profile_collect(ctypes.addressof(kms), ctypes.addressof(kn))
sh.digest.check_weights()
print(json.dumps({"n": n, "ms_per_step": ms, "kernel_ms": kms.value / kn.value, "fixed_us": 1e3 * (ms - kms.value / kn.value),
                  "centroids": len(sh.digest.centroids())}), flush=True)
sh.close()
