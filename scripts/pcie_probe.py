"""Pinned host <-> device copy bandwidth on this box (explains the e2e number)."""
import time, torch, subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("nproc; lscpu | grep -E 'Model name|NUMA|Socket'; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
for mb in (64, 512, 2048):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print("%s %5d MB: %.1f GB/s" % (name, mb, n / dt / 1e9))
# two streams, both directions at once
h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("bidirectional: H2D %.1f GB/s + D2H %.1f GB/s" % (h.numel() / dt / 1e9, h2.numel() / dt / 1e9))
