"""Repeat the attention and scene kernels on many random shapes / seeds and compare every result with a reference
(fp64 torch for the attention, run-to-run determinism + invariants for the scene front end): flushes out rare races."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointnet_refine_b200 import ops, scene as sc
from oracle import scene_oracle as so
dev = torch.device("cuda:0")
rs = np.random.default_rng(0)
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    B = int(rs.integers(1, 400)); N = int(rs.integers(1, 3000))
    g = torch.Generator(device=dev).manual_seed(it)
    qf = (torch.randn(B, 256, 256, device=dev, generator=g) / 8).bfloat16()
    mem = torch.randn(B, N, 256, device=dev, generator=g).bfloat16()
    kp = (mem.float() + 0.5 * torch.randn(B, N, 256, device=dev, generator=g)).bfloat16()
    out = ops.ctx_attention(qf, kp, mem)
    out2 = ops.ctx_attention(qf, kp, mem)
    ref = torch.softmax(qf.double() @ kp.double().transpose(1, 2) * np.log(2.0), dim=-1) @ mem.double()
    err = float((out.double() - ref).abs().max())
    if err > 1e-2 * max(1.0, float(ref.abs().max())) or not torch.equal(out, out2):
        bad += 1; print("attention mismatch", B, N, err, bool(torch.equal(out, out2)))
scene, lines = so.synth_scene(400_000, 40, seed=3, extent=200.0)
d_scene = sc.prepare_scene(torch.from_numpy(scene).to(dev))
first = None
for it in range(20):
    seg = sc.build_segments(d_scene, lines, 1024, 1.0, 2.0, seed=5)
    if first is None:
        first = seg
    elif not (torch.equal(seg.indices, first.indices) and torch.equal(seg.context, first.context) and torch.equal(seg.counts, first.counts)):
        bad += 1; print("scene run-to-run mismatch", it)
for l in (0, 13, 39):
    o = so.build_segment(scene, lines[l], 1024, 1.0, 2.0, 5, l)
    if not np.array_equal(first.indices[l].cpu().numpy(), o[3]):
        bad += 1; print("scene oracle mismatch", l)
print(json.dumps({"stress": "ok" if bad == 0 else "FAILED", "failures": bad}))
