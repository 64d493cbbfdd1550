"""Every tensor-core GEMM shape of the 1024 x 1024 train step on its own: forward / dgrad (K-major pair GEMM,
ops.gemm_bias_act) and wgrad (MN-major split-K, ops.gemm_tn) with the time each takes against the larger of its tensor
time (at the measured cuBLAS burst rate) and its HBM time (operands + result once, measured copy bandwidth).
python tools/train_gemm_shapes.py [points]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointnet_refine_b200 import ops  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024 * 1024
dev = torch.device("cuda:0")
PEAK_TF, PEAK_GB = 1631.0, 6556.5          # MEASURED_PEAKS.json: bf16 burst, HBM copy
layers = [("conv2", 64, 128), ("conv3", 128, 256), ("conv4", 256, 512), ("conv5", 512, 1024), ("fusion", 1984, 1024),
          ("gate2", 64, 1024), ("context_proj", 1024, 256), ("pos_emb2", 256, 256), ("kv_proj", 256, 1536)]


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


rows = []
tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
bound_tot = 0.0
for name, K, N in layers:
    x = torch.randn(P, K, device=dev).bfloat16()
    dy = torch.randn(P, N, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    Kp = max(128, (K + 127) // 128 * 128)          # the backward pads narrow / ragged dgrad outputs to 128 columns
    wt = torch.zeros(Kp, N, device=dev, dtype=torch.bfloat16)
    wt[:K] = w.t()
    xp = torch.zeros(P, Kp, device=dev, dtype=torch.bfloat16)
    xp[:, :K] = x
    flop = 2.0 * P * K * N
    for kind, fn, out_bytes in (("fwd", lambda: ops.gemm_bias_act(x, w, None), P * N * 2),
                                ("dgrad", lambda: ops.gemm_bias_act(dy, wt, None), P * Kp * 2),
                                ("wgrad", lambda: ops.gemm_tn(dy, xp), 0)):
        ms = timed(fn)
        in_bytes = {"fwd": P * K * 2, "dgrad": P * N * 2, "wgrad": P * (Kp + N) * 2}[kind]
        t_tensor, t_hbm = flop / PEAK_TF / 1e9, (in_bytes + out_bytes) / PEAK_GB / 1e6
        bound = max(t_tensor, t_hbm)
        tot[kind] += ms
        bound_tot += bound
        rows.append({"layer": name, "kind": kind, "K": K, "N": N, "ms": round(ms, 3), "tflops": round(flop / ms / 1e9, 1),
                     "gbps": round((in_bytes + out_bytes) / ms / 1e6, 1), "bound": "tensor" if t_tensor > t_hbm else "hbm",
                     "bound_ms": round(bound, 3), "frac": round(bound / ms, 3)})
        print(json.dumps(rows[-1]), flush=True)
    del x, dy, w, wt, xp
print(json.dumps({"points": P, "total_ms": tot, "sum_ms": sum(tot.values()), "sum_bound_ms": bound_tot}))
