"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU fp32.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only
(``python -m oracle.make_golden``): it imports /root/reference/src/model.py,
loads the deterministic synthetic state_dict from oracle/synth.py into the
reference ``LineRefineNet`` (strict), runs it, and stores the (small) outputs.
Weights and inputs are NOT stored: tests regenerate them from the seeds.

Stored per case:
  global_feat (B,2048), argmax (B,1024), gap (B,1024) = top1-top2 of `fused` per
  (segment, channel), fused_sub = fused[:, ::SUB_C, ::SUB_N] in the reference's
  (B,1024,N) layout, fused_csum (B,1024) / fused_psum (B,N) float64 checksums,
  memory_sub / memory_psum likewise for context_proj's output, out (6,B,M,3).
"""
from __future__ import annotations

import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SUB_C, SUB_N = 37, 5

# name: (B, N, weight_seed, input_seed, dist, randomise_bn)
CASES = {
    "b2_n1024": (2, 1024, 0, 1234, "parity", True),
    "b3_n1000_ragged": (3, 1000, 1, 1235, "parity", True),
    "b2_n1031_ragged": (2, 1031, 2, 1236, "parity", True),
    "b5_n37_tiny": (5, 37, 3, 1237, "parity", True),
    "b2_n1_single": (2, 1, 4, 1238, "parity", True),
    "b2_n2048_realistic": (2, 2048, 5, 1239, "realistic", True),
    "b2_n256_defaultbn": (2, 256, 6, 1240, "parity", False),
}
TRAIN_CASE = ("train_b2_n512", 2, 512, 7, 1241)


def main():
    sys.path.insert(0, REF)
    import torch
    from src.model import LineRefineNet  # the reference, unmodified
    from oracle import synth

    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    for name, (B, N, wseed, iseed, dist, rbn) in CASES.items():
        sd = synth.make_state_dict(wseed, rbn)
        ctx, line = synth.make_inputs(B, N, seed=iseed, dist=dist)
        m = LineRefineNet()
        m.load_state_dict(synth.to_torch(sd), strict=True)
        m.eval()
        with torch.no_grad():
            tctx, tline = torch.from_numpy(ctx), torch.from_numpy(line)
            gf, fused = m.context_encoder(tctx.transpose(2, 1))
            vals, idx = torch.max(fused, 2)
            if N > 1:
                top2 = torch.topk(fused, 2, dim=2).values
                gap = (top2[..., 0] - top2[..., 1])
            else:
                gap = torch.full_like(vals, float("inf"))
            memory = m.context_proj(fused.transpose(2, 1))
            out = m(tctx, tline)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            meta=np.array([B, N, wseed, iseed, int(rbn), SUB_C, SUB_N], np.int64), dist=np.array(dist),
            global_feat=gf.numpy(), argmax=idx.numpy(), gap=gap.numpy(),
            fused_sub=fused[:, ::SUB_C, ::SUB_N].contiguous().numpy(),
            fused_csum=fused.double().sum(2).numpy(), fused_psum=fused.double().sum(1).numpy(),
            memory_sub=memory[:, ::SUB_N, ::SUB_C].contiguous().numpy(),
            memory_psum=memory.double().sum(2).numpy(),
            out=out.numpy(),
        )
        print(name, "gf.sum", float(gf.double().sum()), "out.sum", float(out.double().sum()))

    name, B, N, wseed, iseed = TRAIN_CASE
    sd = synth.make_state_dict(wseed, True)
    ctx, _ = synth.make_inputs(B, N, seed=iseed)
    m = LineRefineNet()
    m.load_state_dict(synth.to_torch(sd), strict=True)
    m.train()
    with torch.no_grad():
        gf, fused = m.context_encoder(torch.from_numpy(ctx).transpose(2, 1))
    new_sd = m.state_dict()
    stats = {k.replace(".", "__"): v.numpy() for k, v in new_sd.items()
             if k.startswith("context_encoder.") and ("running_" in k or "num_batches" in k)}
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=np.array([B, N, wseed, iseed, 1, SUB_C, SUB_N], np.int64),
        global_feat=gf.numpy(), fused_sub=fused[:, ::SUB_C, ::SUB_N].contiguous().numpy(),
        fused_csum=fused.double().sum(2).numpy(), **stats)
    print(name, "gf.sum", float(gf.double().sum()))


if __name__ == "__main__":
    main()
