"""Many digests x many repeated large updates: every digest must end up identical."""
import sys, collections
sys.path.insert(0, "/root/repo")
import torch
import crick_b200 as cb
from crick_b200._lib import lib
torch.manual_seed(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
x = torch.randn(n, dtype=torch.float64, device="cuda").exp_()
w = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
torch.cuda.synchronize()
seen = collections.Counter()
for rep in range(reps):
    t = cb.TDigest(100)
    for _ in range(20):
        t.update(x, w, mode="parallel", check_w=False)
    seen[(float(t.quantile(0.5)), t.size(), len(t.centroids()))] += 1
for k, v in seen.items():
    print(v, k)
print("distinct:", len(seen))
