"""Train-step timing (BASELINE.json configs[3] shape class): LineRefineNet forward + L1 deep-supervision loss +
backward + Adam, native encoder path vs the stock-PyTorch formulation of the encoder."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnet_refine_b200 as prb
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
only_encoder = "--encoder" in sys.argv
dev = torch.device("cuda:0")
res = {}
for native in (True, False):
    torch.manual_seed(0)
    m = prb.LineRefineNet().to(dev).train()
    m.context_encoder.native_training = native
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    ctx = torch.randn(B, N, 4, device=dev); line = torch.randn(B, 32, 3, device=dev); tgt = torch.randn(B, 32, 3, device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        if only_encoder:
            gf, fused = m.context_encoder(ctx.transpose(2, 1))
            loss = fused.mean() + gf.mean()
        else:
            out = m(ctx, line)
            loss = sum(torch.nn.functional.l1_loss(out[l], tgt) for l in range(out.shape[0])) / out.shape[0]
        loss.backward()
        opt.step()
        return loss
    try:
        for _ in range(2): step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): l = step()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        res["native" if native else "torch"] = {"ms_per_step": round(dt * 1e3, 1), "segments_per_s": round(B / dt, 1), "loss": float(l.detach()),
                                                 "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
    except torch.cuda.OutOfMemoryError as e:
        res["native" if native else "torch"] = "OOM"
    del m, opt; torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
print(json.dumps({"B": B, "N": N, "encoder_only": only_encoder, **res}))
