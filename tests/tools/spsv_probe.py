"""SpaceSaving counting pass probe: n Zipf(1.1)-like int64 items on the device, capacity 1000;
whole-call time of update / update+moments fused / SummaryStats alone.  CRICK_HH_V=1 forces the
round-1 counting kernel.   python tests/tools/spsv_probe.py [n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
dx = cb.DeviceArray((n,), "i8")
lib.crick_synth_fill(dx.ptr, None, n, 5, 0, 4, None)
lib.crick_device_synchronize()

def timed(fn, reps=5):
    fn(); lib.crick_device_synchronize()
    best = 1e9
    for _ in range(reps):
        lib.crick_device_synchronize(); t0 = time.perf_counter(); fn(); lib.crick_device_synchronize()
        best = min(best, time.perf_counter() - t0)
    return best

out = {"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("CRICK_")}}
keep = {}
def spsv():
    s = cb.SpaceSaving(1000, "i8"); s.update(dx); keep["s"] = s
def fused():
    s = cb.SpaceSaving(1000, "i8"); st = cb.SummaryStats(); s.update(dx, stats=st); keep["f"] = (s, st)
def stats():
    st = cb.SummaryStats(); st.update(dx); keep["st"] = st
for name, fn in (("spsv", spsv), ("fused", fused), ("stats", stats)):
    t = timed(fn)
    out[name] = {"ms": t * 1e3, "Gvalues_s": n / t / 1e9, "frac_of_6542": 8 * n / t / 1e9 / 6542.4}
c = keep["s"].counters()
out["max_error"] = int(c["error"].max()); out["min_count"] = int(c["count"].min())
out["fused_equal"] = keep["f"][0].counters().tobytes() == c.tobytes()
out["stats_count"] = keep["f"][1].count()
print(json.dumps(out), flush=True)
