"""Compare the native backward's workspace intermediates with an fp32 torch restatement (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import pointnet_refine_b200 as prb
from pointnet_refine_b200 import train_ops
from oracle import synth
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, N = 4, 600
P = B * N; Pp = (P + 255) // 256 * 256
m = prb.LineRefineNet().to(dev); m.load_state_dict(synth.to_torch(synth.make_state_dict(7)))
enc = m.context_encoder.train()
enc.native_training = True
ctx = torch.from_numpy(synth.make_inputs(B, N, seed=1241)[0]).to(dev)
R = torch.randn(B, 1024, N, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
# ---- native
holder = {}
orig = train_ops.EncoderTrainFn.backward
def patched(c, d):
    holder["ws"] = c.ws
    return orig(c, d)
train_ops.EncoderTrainFn.backward = staticmethod(patched)
gf, fused = enc(ctx.transpose(2, 1))
((fused * R).sum() / P).backward()
torch.cuda.synchronize()
ws = holder["ws"]
al = lambda v: (v + 1023) // 1024 * 1024
off = 0; lay = {}
def take(name, nbytes):
    global off
    lay[name] = off; off = al(off + nbytes)
take("X", Pp * 2048 * 2); take("U", Pp * 3008 * 2); take("Z", Pp * 1024 * 2); take("stats", (8 * 3008 + 1024) * 4)
chan = [4, 64, 128, 256, 512, 1024]
for k in range(2, 6): take(f"w{k}", chan[k] * chan[k - 1] * 2)
take("wf", 1024 * 2048 * 2); take("wg2", 1024 * 64 * 2)
for k in range(2, 6): take(f"wt{k}", max(chan[k - 1], 128) * chan[k] * 2)
take("wft", 2048 * 1024 * 2); take("wg2t", 128 * 1024 * 2)
take("XT", 2048 * Pp * 2); take("dA", Pp * 2048 * 2); take("dB", Pp * 512 * 2); take("dU", Pp * 1024 * 2)
take("dUT", 1024 * Pp * 2); take("dZ", Pp * 1024 * 2); take("dZT", 1024 * Pp * 2); take("dHp", Pp * 128 * 2); take("gW", 1024 * 2048 * 4)
def bf(name, rows, cols): return ws[lay[name]: lay[name] + rows * cols * 2].view(torch.bfloat16).view(rows, cols).float()
X = bf("X", Pp, 2048)[:P]; U = bf("U", Pp, 3008)[:P]; Z = bf("Z", Pp, 1024)[:P]
XT = bf("XT", 2048, Pp); dA = bf("dA", Pp, 2048)[:P]; dZ = bf("dZ", Pp, 1024)[:P]; dZT = bf("dZT", 1024, Pp)
print("XT == X^T:", torch.equal(XT[:, :P], X.T.contiguous()), " pad zero:", float(XT[:, P:].abs().max()))
print("dZT == dZ^T:", torch.equal(dZT[:, :P], dZ.T.contiguous()), " pad zero:", float(dZT[:, P:].abs().max()))
# ---- fp32 restatement
sd = {k: v.detach() for k, v in enc.state_dict().items()}
x = ctx.reshape(P, 4)
feats = []; h = x
Us = []
for k in range(1, 6):
    u = h @ sd[f"conv{k}.weight"].squeeze(-1).T + sd[f"conv{k}.bias"]
    mu, var = u.mean(0), u.var(0, unbiased=False)
    xh = (u - mu) / torch.sqrt(var + 1e-5)
    h = torch.relu(xh * sd[f"bn{k}.weight"] + sd[f"bn{k}.bias"])
    feats.append(h); Us.append(u)
cat = torch.cat(feats, 1)
uf = cat @ sd["fusion.0.weight"].squeeze(-1).T + sd["fusion.0.bias"]
mu, var = uf.mean(0), uf.var(0, unbiased=False); rstd = 1 / torch.sqrt(var + 1e-5)
xh = (uf - mu) * rstd
y0 = xh * sd["fusion.1.weight"] + sd["fusion.1.bias"]
hg = torch.relu(x[:, 3:4] @ sd["intensity_gate.0.weight"].squeeze(-1).T + sd["intensity_gate.0.bias"])
z = hg @ sd["intensity_gate.2.weight"].squeeze(-1).T + sd["intensity_gate.2.bias"]
g = torch.sigmoid(z)
print("X feats err", float((X[:, :1984] - cat).abs().max()), "range", float(cat.abs().max()), "| H err", float((X[:, 1984:] - hg).abs().max()))
print("Uf err", float((U[:, 1984:] - uf).abs().max()), "range", float(uf.abs().max()), "| Z err", float((Z - z).abs().max()), "range", float(z.abs().max()))
dF = (R / P).permute(0, 2, 1).reshape(P, 1024)
dY = dF * (0.5 + 0.5 * g) * (y0 > 0)
dZr = dF * torch.relu(y0) * 0.5 * g * (1 - g)
print("dZ err", float((dZ - dZr).abs().max()), "range", float(dZr.abs().max()))
S1, S2 = dY.sum(0), (dY * xh).sum(0)
dUf = sd["fusion.1.weight"] * rstd * (dY - S1 / P - xh * S2 / P)
dAr = dUf @ sd["fusion.0.weight"].squeeze(-1)
print("dA err", float((dA[:, :1984] - dAr).abs().max()), "range", float(dAr.abs().max()))
dWf_ref = dUf.T @ cat
print("dWf native-vs-restatement err", float((enc.fusion[0].weight.grad.squeeze(-1) - dWf_ref).abs().max()), "range", float(dWf_ref.abs().max()))
dWg2_ref = dZr.T @ hg
print("dWg2 err", float((enc.intensity_gate[2].weight.grad.squeeze(-1) - dWg2_ref).abs().max()), "range", float(dWg2_ref.abs().max()))
# emulate with the native's own (bf16) intermediates
dWg2_emul = dZ.T @ X[:, 1984:]
print("dWg2 vs emulation from native dZ/H:", float((enc.intensity_gate[2].weight.grad.squeeze(-1) - dWg2_emul).abs().max()))
rel = lambda a, b: float((a - b).norm() / b.norm())
print("rel-L2: X", rel(X[:, :1984], cat), "Uf", rel(U[:, 1984:], uf), "dZ", rel(dZ, dZr), "dA", rel(dA[:, :1984], dAr),
      "dWf", rel(enc.fusion[0].weight.grad.squeeze(-1), dWf_ref), "dWg2", rel(enc.intensity_gate[2].weight.grad.squeeze(-1), dWg2_ref))
# emulate dU_f from the native's own bf16 buffers to separate arithmetic from rounding
Ufn = U[:, 1984:]
mun, varn = Ufn.mean(0), Ufn.var(0, unbiased=False); rstdn = 1 / torch.sqrt(varn + 1e-5)
xhn = (Ufn - mun) * rstdn
y0n = xhn * sd["fusion.1.weight"] + sd["fusion.1.bias"]
gn = torch.sigmoid(Z)
dYn = (dF * (0.5 + 0.5 * gn) * (y0n > 0)).bfloat16().float()
S1n, S2n = dYn.sum(0), (dYn * xhn).sum(0)
dUn = (sd["fusion.1.weight"] * rstdn * (dYn - S1n / P - xhn * S2n / P)).bfloat16().float()
dAn = dUn @ sd["fusion.0.weight"].squeeze(-1).bfloat16().float()
print("dA vs emulation of the native arithmetic: rel-L2", rel(dA[:, :1984], dAn), " | emulation vs fp32 restatement:", rel(dAn, dAr))
print("dU: emulated-bf16 vs fp32 restatement rel-L2", rel(dUn, dUf), " dY:", rel(dYn, dY))
print("fraction of sign flips in ReLU mask:", float(((y0n > 0) != (y0 > 0)).float().mean()))
