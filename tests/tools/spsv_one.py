import sys
sys.path.insert(0, "/root/repo")
import torch, crick_b200 as cb
n = int(float(sys.argv[1]))
u = torch.rand(n, dtype=torch.float64, device="cuda")
x = u.pow_(-10.0).clamp_(max=4.0e18).to(torch.int64)
torch.cuda.synchronize()
for _ in range(2):
    s = cb.SpaceSaving(1000, "i8")
    s.update(x, mode=3)
    print(s.size())
