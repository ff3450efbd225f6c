import sys, time
sys.path.insert(0, "/root/repo")
import crick_b200 as cb
from crick_b200._lib import lib
t0 = time.perf_counter()
ds = [cb.TDigest(100) for _ in range(2000)]
t1 = time.perf_counter()
del ds
lib.crick_device_synchronize()
t2 = time.perf_counter()
print("TDigest(): %.1f us to create, %.1f us to free" % ((t1 - t0) / 2000 * 1e6, (t2 - t1) / 2000 * 1e6))
