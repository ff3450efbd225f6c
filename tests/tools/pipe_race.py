"""Run-to-run determinism of the pipelined update forms: md5 of the centroids after 3 x 5 updates of
4 Mi values, several repetitions per (overlap, reserve) setting, plus the plain call."""
import hashlib, json, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import crick_b200 as cb
from crick_b200._lib import lib

rng = np.random.default_rng(11)
n = 1 << 22
dev = [(cb.DeviceArray.from_numpy(rng.lognormal(0, 1 + 0.2 * i, n)), cb.DeviceArray.from_numpy(rng.uniform(0.5, 2.0, n)))
       for i in range(5)]
s1, s2 = lib.crick_stream_create(), lib.crick_stream_create()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
out = {}
for name, overlap, reserve in (("plain", None, 0), ("pipe_r0", 0, 0), ("pipe_r2", 0, 2), ("pipe_r8", 0, 8), ("overlap_r8", 1, 8)):
    if only and name not in only:
        continue
    if overlap is not None:
        lib.crick_pipeline_overlap(overlap); lib.crick_pipeline_reserve_sms(reserve)
    hs = []
    for _ in range(reps):
        c = cb.TDigest(100, mode="parallel")
        for rep in range(3):
            for dx, dw in dev:
                if overlap is None:
                    c.update(dx, dw, mode="parallel", stream=s1, check_w="deferred")
                else:
                    c.update(dx, dw, stream=s1, tail_stream=s2)
        c.check_weights()
        hs.append(hashlib.md5(c.centroids().tobytes()).hexdigest()[:8])
    out[name] = hs
print(json.dumps(out))
