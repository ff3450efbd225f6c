import sys, time
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib
torch.manual_seed(0)
n = 100_000_000
x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
torch.cuda.synchronize()
true_med = float(x[:10_000_000].median())
for trial in range(6):
    t = cb.TDigest(100)
    for rep in range(1 + 3 * (trial % 2)):
        t.update(x, w, mode="parallel", check_w=False)
    lib.crick_device_synchronize()
    c = t.centroids()
    print("trial", trial, "q50", float(t.quantile(0.5)), "approx true", true_med, "size/n", t.size() / n, "ncen", len(c), "min", t.min(), "max", t.max(), flush=True)
