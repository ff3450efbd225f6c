"""Where does a mid-size update call spend its time: host issue time vs GPU time."""
import sys, time, ctypes
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
torch.cuda.synchronize()
t = cb.TDigest(100)
for _ in range(5):
    t.update(x, w, mode="parallel", check_w=False)
lib.crick_device_synchronize()
reps = 200
t0 = time.perf_counter()
for _ in range(reps):
    t.update(x, w, mode="parallel", check_w=False)
t1 = time.perf_counter()
lib.crick_device_synchronize()
t2 = time.perf_counter()
print("wrapper : issue %.1f us/call, total %.1f us/call" % ((t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6))
xp, wp = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr())
t0 = time.perf_counter()
for _ in range(reps):
    lib.crick_tdigest_update(t._h, xp, wp, 1.0, n, 1, 1, 2, None)
t1 = time.perf_counter()
lib.crick_device_synchronize()
t2 = time.perf_counter()
print("C ABI   : issue %.1f us/call, total %.1f us/call" % ((t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6))
