"""Run by tests/test_gpu_variants.py in a subprocess with a kernel-selection environment variable set:
checks the bf16 encoder against the reference fixtures on the selected code path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pointnet_refine_b200 as prb
from oracle import synth
from tests.golden_util import load_case

dev = torch.device("cuda:0")
for name in ("b2_n1024", "b3_n1000_ragged", "b5_n37_tiny"):
    g, sd, ctx, line, (sc, sn) = load_case(name)
    m = prb.LineRefineNet().to(dev).eval()
    m.load_state_dict(synth.to_torch(sd), strict=True)
    m.precision = "bf16"
    with torch.no_grad():
        out = m.context_encoder.run_native(torch.from_numpy(ctx).to(dev), pool=True, fused=True, memory=True)
        out2 = m.context_encoder.run_native(torch.from_numpy(ctx).to(dev), pool=True, argmax=True)
    rng = max(1.0, float(np.abs(g["global_feat"]).max()))
    gf, fused, mem = out["global_feat"].cpu().numpy(), out["fused"].cpu().numpy(), out["memory"].cpu().numpy()
    assert np.abs(gf - g["global_feat"]).max() <= 1e-2 * rng, name
    assert np.abs(fused[:, ::sc, ::sn] - g["fused_sub"]).max() <= 1e-2 * rng, name
    assert np.abs(mem[:, ::sn, ::sc] - g["memory_sub"]).max() <= 1e-2 * max(1.0, float(np.abs(g["memory_sub"]).max())), name
    assert np.array_equal(gf[:, :1024], fused.max(axis=2)), name
    assert np.array_equal(out2["global_feat"][:, :1024].cpu().numpy(), gf[:, :1024]), name
    assert np.array_equal(out2["argmax"].cpu().numpy(), fused.argmax(axis=2)), name
    with torch.no_grad():   # whole forward on the selected decoder path (LRN_FAST_DECODER / LRN_CTX_ATTN)
        full = m(torch.from_numpy(ctx).to(dev), torch.from_numpy(line).to(dev)).cpu().numpy()
    assert np.abs(full - g["out"]).max() <= 5e-2 * max(1.0, float(np.abs(g["out"]).max())), name
print("variant ok", {k: v for k, v in os.environ.items() if k.startswith("LRN_")})
