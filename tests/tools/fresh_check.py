import sys
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
torch.manual_seed(0)
n = 100_000_000
# same RNG consumption as latency_probe before this size
for m in (10_000, 100_000, 1_000_000, 10_000_000):
    torch.randn(m, dtype=torch.float64, device="cuda"); torch.rand(m, dtype=torch.float64, device="cuda")
x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
torch.cuda.synchronize()
t = cb.TDigest(100)
for _ in range(53):
    t.update(x, w, mode="parallel", check_w=False)
print("%.6f %.3f %d" % (float(t.quantile(0.5)), t.size(), len(t.centroids())))
