"""torch.profiler breakdown of one full native train step (LineRefineNet fwd + loss + bwd + Adam)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import pointnet_refine_b200 as prb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = prb.LineRefineNet().to(dev).train()
m.context_encoder.native_training = True
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
ctx = torch.randn(B, N, 4, device=dev); line = torch.randn(B, 32, 3, device=dev); tgt = torch.randn(B, 32, 3, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    out = m(ctx, line)
    loss = sum(torch.nn.functional.l1_loss(out[l], tgt) for l in range(6)) / 6
    loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
print(f"total device time {tot/1e3:.1f} ms for {B}x{N}")
for e in ev[:60]:
    print(f"{e.device_time_total/1e3:8.2f} ms {100*e.device_time_total/tot:5.1f}% x{e.count:<4d} {e.key[:100]}")
