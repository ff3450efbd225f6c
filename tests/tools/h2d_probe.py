import torch, time, os
print(open('/proc/cpuinfo').read().count('processor\t'), 'cpus')
os.system("lscpu | grep -i -E 'numa|socket|model name' ; nvidia-smi topo -m | head -8; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('H2D pinned 1 GiB x4: %.1f GB/s' % (4 * n / dt / 1e9))
# two streams concurrently
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('H2D two streams: %.1f GB/s' % (8 * n / dt / 1e9))
