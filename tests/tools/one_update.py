"""A few parallel updates of N device-resident values (for ncu launch lists of the setup kernels)."""
import sys
sys.path.insert(0, "/root/repo")
import torch, crick_b200 as cb
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
torch.cuda.synchronize()
t = cb.TDigest(100)
for _ in range(4):
    t.update(x, w, mode="parallel", check_w=False)
print(t.quantile(0.5))
