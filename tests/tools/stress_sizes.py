"""latency_probe's sequence (growing sizes in one process, 53 repeated updates per digest),
repeated; prints every result that differs from the first pass."""
import sys
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 10
first = {}
bad = 0
for p in range(passes):
    torch.manual_seed(0)
    for n in (10_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
        x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
        w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
        torch.cuda.synchronize()
        for weighted in (True, False):
            t = cb.TDigest(100)
            args = (x, w) if weighted else (x,)
            for _ in range(53):
                t.update(*args, mode="parallel", check_w=False)
            lib.crick_device_synchronize()
            key = (n, weighted)
            val = (float(t.quantile(0.5)), t.size(), len(t.centroids()))
            if key not in first:
                first[key] = val
            elif first[key] != val:
                bad += 1
                print("MISMATCH pass", p, key, val, "expected", first[key], flush=True)
        del x, w
print("passes", passes, "mismatches", bad)
