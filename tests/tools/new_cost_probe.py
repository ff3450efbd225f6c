"""Cost of creating / freeing digests (dask: one digest per chunk)."""
import sys, time
sys.path.insert(0, "/root/repo")
import crick_b200 as cb
from crick_b200._lib import lib
N = 2000
t0 = time.perf_counter()
hs = [lib.crick_tdigest_new(100.0) for _ in range(N)]
t1 = time.perf_counter()
for h in hs: lib.crick_tdigest_free(h)
t2 = time.perf_counter()
print("C ABI  : %.1f us to create, %.1f us to free (2000 live)" % ((t1 - t0) / N * 1e6, (t2 - t1) / N * 1e6))
t0 = time.perf_counter()
for _ in range(N):
    h = lib.crick_tdigest_new(100.0); lib.crick_tdigest_free(h)
t1 = time.perf_counter()
print("C ABI  : %.1f us per create+free pair (one live)" % ((t1 - t0) / N * 1e6))
t0 = time.perf_counter()
ds = [cb.TDigest(100) for _ in range(N)]
t1 = time.perf_counter()
del ds
lib.crick_device_synchronize()
t2 = time.perf_counter()
print("Python : %.1f us to create, %.1f us to free (2000 live)" % ((t1 - t0) / N * 1e6, (t2 - t1) / N * 1e6))
