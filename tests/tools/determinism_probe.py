"""Repeat the same parallel updates many times: the centroid bytes must never change (the sums
are ordered, the setup chain is launched with programmatic dependencies)."""
import sys, hashlib
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import crick_b200 as cb

torch.manual_seed(1)
bad = 0
for n in (20_000, 300_000, 5_000_000):
    x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
    w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
    torch.cuda.synchronize()
    seen = set()
    for rep in range(300):
        t = cb.TDigest(100)
        t.update(x, w, mode="parallel")
        t.update(x, mode="parallel")
        seen.add(hashlib.sha1(t.centroids().tobytes()).hexdigest())
    print(n, "distinct results:", len(seen))
    bad += len(seen) != 1
sys.exit(1 if bad else 0)
