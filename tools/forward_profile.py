"""Kernel-level breakdown of the whole LineRefineNet eval forward (torch profiler, CUDA activity)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import pointnet_refine_b200 as prb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
m = prb.LineRefineNet().to(dev).eval()
ctx = torch.randn(B, N, 4, device=dev); line = torch.randn(B, 32, 3, device=dev)
with torch.no_grad():
    for _ in range(3): m(ctx, line)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3): m(ctx, line)
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device ms/forward {tot:.2f}")
for k, t, n in rows[:28]:
    print(f"{t:8.3f} ms  x{n:<4d} {k[:110]}")
