import torch, time, numpy as np
n = 1_000_000
src = np.random.rand(n)
h = torch.empty(8 * 1024 * 1024, dtype=torch.float64).pin_memory()    # 64 MB pinned like a staging slot
d = torch.empty(n, dtype=torch.float64, device="cuda")
hv = h.numpy()
for mode in ("no cpu write", "cpu write first", "cpu write first, 8 MB only pinned"):
    if mode.endswith("only pinned"):
        h = torch.empty(n, dtype=torch.float64).pin_memory(); hv = h.numpy()
    ts = []
    for rep in range(8):
        if mode != "no cpu write":
            hv[:n] = src
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d.copy_(h[:n], non_blocking=True); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    print("%-36s H2D 8 MB: %s us" % (mode, " ".join("%.0f" % t for t in ts)))
