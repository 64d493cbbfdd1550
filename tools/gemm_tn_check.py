import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet_refine_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
for (M, N, K) in [(128, 128, 64), (256, 256, 128), (1024, 2048, 2560), (128, 128, 25600), (512, 256, 4096), (64, 128, 640)]:
    at = torch.randn(K, M, device=dev).bfloat16(); bt = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()
    out = ops.gemm_tn(at, bt)
    ref = at.double().T @ bt.double()
    print(f"gemm_tn M={M} N={N} K={K}: max err {(out.double() - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.2f})")
# strided views (columns of wider buffers), as the backward uses them
buf = torch.randn(2560, 3008, device=dev).bfloat16(); xb = (torch.randn(2560, 2048, device=dev) / 50).bfloat16()
out = ops.gemm_tn(buf[:, 960:1984], xb[:, 448:960])
ref = buf[:, 960:1984].double().T @ xb[:, 448:960].double()
print("strided views: max err", (out.double() - ref).abs().max().item(), "ref max", ref.abs().max().item())
