"""TDigest.update from host memory at dask-chunk sizes: pageable NumPy vs pinned vs device."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib
rng = np.random.default_rng(0)
def timed(fn, reps=5):
    fn(); fn(); lib.crick_device_synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    lib.crick_device_synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for n in (10**5, 10**6, 4 * 10**6, 10**7, 3 * 10**7):
    x = rng.lognormal(size=n)
    px = cb.PinnedArray(n); px.array[:] = x
    dx = cb.DeviceArray.from_numpy(x)
    t = cb.TDigest(100)
    a = timed(lambda: t.update(x)); b = timed(lambda: t.update(px.array)); c = timed(lambda: t.update(dx))
    print("n=%9d  pageable %7.3f ms (%5.1f GB/s)   pinned %7.3f ms (%5.1f GB/s)   device %6.3f ms" % (
        n, a, 8 * n / a / 1e6, b, 8 * n / b / 1e6, c), flush=True)
