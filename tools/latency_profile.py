"""Per-kernel device time of one B = 1 whole-scene style forward (run under `ncu --metrics gpu__time_duration.sum`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnet_refine_b200 as prb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
m = prb.LineRefineNet().to(dev).eval()
ctx = torch.randn(B, N, 4, device=dev); line = torch.randn(B, 32, 3, device=dev)
with torch.no_grad():
    for _ in range(3):
        m(ctx, line)
torch.cuda.synchronize()
