"""Throughput probe of the hot kernel (k_bin_accumulate / k_bin_accumulate2): weighted and
unweighted TDigest.update over n device-resident lognormal values, kernel time from the library's
own CUDA events.  Environment: CRICK_ACC_V=1 forces the round-1 kernel, CRICK_ACC2_T / _R the
table / ring shape of the round-2 one.   python tests/tools/acc_probe.py [n] [kind]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crick_b200 as cb
from crick_b200._lib import lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
lib.crick_synth_fill(dx.ptr, dw.ptr, n, 1234, 0, kind, None)
lib.crick_device_synchronize()
out = {"n": n, "kind": kind, "env": {k: v for k, v in os.environ.items() if k.startswith("CRICK_")}}
for name, w in (("weighted", dw), ("unweighted", None)):
    t = cb.TDigest(100, mode="parallel")
    for _ in range(3):
        t.update(dx, w, check_w=False) if w is not None else t.update(dx)
    lib.crick_device_synchronize()
    lib.crick_profile_enable(1)
    for _ in range(10):
        t.update(dx, w, check_w=False) if w is not None else t.update(dx)
    lib.crick_device_synchronize()
    lib.crick_profile_enable(0)
    ms, k = ctypes.c_double(0), ctypes.c_int64(0)
    lib.crick_profile_collect(ctypes.addressof(ms), ctypes.addressof(k))
    per = ms.value / max(k.value, 1)
    bpv = 16 if w is not None else 8
    out[name] = {"kernel_ms": per, "Gvalues_s": n / per / 1e6, "GBs": bpv * n / per / 1e6,
                 "frac_of_6542": bpv * n / per / 1e6 / 6542.4, "size": t.size(), "ncen": len(t.centroids())}
print(json.dumps(out), flush=True)
