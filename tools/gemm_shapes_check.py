import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet_refine_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (M, N, K, od) in [(1024, 128, 2560, torch.float32), (1024, 2048, 2560, torch.float32), (2400, 2048, 1024, torch.bfloat16),
                      (128, 128, 2560, torch.float32), (256, 128, 2560, torch.float32), (512, 256, 25600, torch.float32),
                      (2400, 128, 1024, torch.bfloat16), (2400, 512, 1024, torch.bfloat16)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    out = ops.gemm_bias_act(a, w, None, relu=False, out_dtype=od)
    ref = a.double() @ w.double().T
    err = (out.double() - ref).abs().max().item()
    print(f"M={M} N={N} K={K} out={od}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})")
# strided A (lda > K) as used by the backward: A = view into a wider buffer
buf = torch.randn(2400, 1024, device=dev).bfloat16()
w = (torch.randn(128, 256, device=dev) / 16).bfloat16()
