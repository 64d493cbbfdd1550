"""Diagnostic: native train-mode encoder (forward + backward) vs the stock-PyTorch formulation in fp32."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pointnet_refine_b200 as prb
from oracle import synth
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 600)
m = prb.LineRefineNet().to(dev)
m.load_state_dict(synth.to_torch(synth.make_state_dict(7)))
enc = m.context_encoder.train()
enc.native_training = True
ref = copy.deepcopy(enc).train()
ref.native_training = False
ctx = torch.from_numpy(synth.make_inputs(B, N, seed=1241)[0]).to(dev)
g = torch.Generator(device=dev).manual_seed(3)
R = torch.randn(B, 1024, N, device=dev, generator=g)
R2 = torch.randn(B, 2048, device=dev, generator=g)

def run(e):
    gf, fused = e(ctx.transpose(2, 1))
    loss = (fused * R).sum() / (B * N) + (gf * R2).sum() / B
    loss.backward()
    return gf.detach(), fused.detach(), loss.detach()

gf_r, fz_r, l_r = run(ref)
gf_n, fz_n, l_n = run(enc)
torch.cuda.synchronize()
rng = fz_r.abs().max().item()
print(f"fused max-abs err {(fz_n - fz_r).abs().max().item():.3e} (range {rng:.3f})  global_feat err {(gf_n - gf_r).abs().max().item():.3e}  loss {l_n.item():.5f} vs {l_r.item():.5f}")
for (n, p), (_, q) in zip(enc.named_buffers(), ref.named_buffers()):
    if p.dtype.is_floating_point:
        e = (p - q).abs().max().item() / max(q.abs().max().item(), 1e-6)
        if e > 1e-3: print(f"  buffer {n}: rel err {e:.3e}")
    else:
        assert int(p) == int(q), n
worst = 0
for (n, p), (_, q) in zip(enc.named_parameters(), ref.named_parameters()):
    e = (p.grad - q.grad).abs().max().item()
    s = q.grad.abs().max().item()
    worst = max(worst, e / max(s, 1e-12) if s > 1e-6 else 0)
    print(f"  grad {n:32s} max-abs err {e:.3e}  ref max {s:.3e}  rel {e / max(s, 1e-12):.3e}")
print("worst relative grad error (tensors with non-negligible grads):", worst)
