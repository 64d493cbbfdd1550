"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): eval encoder (all outputs, both
tiers), full forward, heads, train forward/backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnet_refine_b200 as prb
from oracle import synth
dev = torch.device("cuda:0")
m = prb.LineRefineNet().to(dev).eval()
m.load_state_dict(synth.to_torch(synth.make_state_dict(0)))
ctx, line = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(3, 300, seed=1))
for prec in ("bf16", "tf32"):
    m.precision = prec
    with torch.no_grad():
        m.context_encoder.run_native(ctx, pool=True, argmax=True, fused=True, memory=True)
        m.context_encoder.run_native(ctx, pool=True)
        out = m(ctx, line)
m.precision = "bf16"
m.train()
out = m(ctx, line)
out.abs().mean().backward()
torch.cuda.synchronize()
print("sanitize target ok", float(out.abs().mean()))
