"""Quick on-GPU diagnostic (not a test): GEMM unit check, encoder vs oracle, timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pointnet_refine_b200 as prb
from pointnet_refine_b200 import ops
from oracle import synth, lrn_oracle as orc

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)
torch.manual_seed(0)

def gemm_case(M, N, K, dtype, relu, out_dtype):
    a = torch.randn(M, K, device=dev).to(dtype)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dtype)
    b = torch.randn(N, device=dev)
    out = ops.gemm_bias_act(a, w, b, relu=relu, out_dtype=out_dtype)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T + b.double()
    if relu: ref = ref.clamp_min(0)
    err = (out.double() - ref).abs().max().item()
    print(f"gemm M={M} N={N} K={K} {dtype} relu={relu} out={out_dtype}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})", flush=True)
    return err

for (M, N, K) in [(128, 128, 64), (128, 256, 64), (300, 128, 128), (1000, 256, 512), (4096, 1024, 512), (20000, 512, 256)]:
    gemm_case(M, N, K, torch.bfloat16, True, torch.float32)
gemm_case(777, 256, 256, torch.bfloat16, False, torch.bfloat16)
for (M, N, K) in [(128, 128, 32), (1000, 256, 512), (4096, 1024, 512)]:
    gemm_case(M, N, K, torch.float32, True, torch.float32)

for prec in ("bf16", "tf32"):
    for (B, N) in [(2, 1024), (3, 1000), (5, 37)]:
        sd = synth.make_state_dict(0)
        ctx, line = synth.make_inputs(B, N, seed=1234)
        gf_o, fused_o, arg_o = orc.encoder_forward(sd, ctx)
        mem_o = orc.context_memory(sd, fused_o)
        m = prb.LineRefineNet().to(dev).eval()
        m.load_state_dict(synth.to_torch(sd))
        m.precision = prec
        with torch.no_grad():
            out = m.context_encoder.run_native(torch.from_numpy(ctx).to(dev), pool=True, argmax=True, fused=True, memory=True)
            torch.cuda.synchronize()
        gf = out["global_feat"].cpu().numpy(); fz = out["fused"].cpu().numpy(); am = out["argmax"].cpu().numpy(); mem = out["memory"].cpu().numpy()
        rng_ = np.abs(gf_o).max()
        print(f"[{prec}] B={B} N={N}: gf max-abs {np.abs(gf-gf_o).max():.3e} (range {rng_:.3f}) "
              f"fused {np.abs(fz - fused_o.transpose(0,2,1)).max():.3e} mem {np.abs(mem-mem_o).max():.3e} (range {np.abs(mem_o).max():.2f}) "
              f"argmax agree {(am==arg_o).mean():.4f}", flush=True)
        with torch.no_grad():
            o = m(torch.from_numpy(ctx).to(dev), torch.from_numpy(line).to(dev)).cpu().numpy()
        oo = orc.line_refine_forward(sd, ctx, line)
        print(f"    full forward max-abs {np.abs(o-oo).max():.3e} (range {np.abs(oo).max():.2f})", flush=True)

# timing at a mid size
m = prb.LineRefineNet().to(dev).eval()
for prec in ("bf16", "tf32"):
    m.precision = prec
    B, N = 1024, 4096
    ctx = torch.randn(B, N, 4, device=dev)
    with torch.no_grad():
        for _ in range(2): m.context_encoder.run_native(ctx, pool=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3): m.context_encoder.run_native(ctx, pool=True)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    pts = B * N / ms * 1e3
    print(f"[{prec}] encoder pool-only {B}x{N}: {ms:.2f} ms  {pts/1e6:.1f} Mpts/s  {pts*5.587584e6/1e12:.1f} TFLOP/s", flush=True)
