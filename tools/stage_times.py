"""Per-stage device times of the encoder (events around every launch) -- tuning aid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnet_refine_b200 as prb
from pointnet_refine_b200 import ops
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
m = prb.LineRefineNet().to(dev).eval()
m.load_state_dict(synth.to_torch(synth.make_state_dict(0)))
m.precision = prec
m.context_encoder.chunk_rows = chunk
ctx = torch.randn(B, N, 4, device=dev)
with torch.no_grad():
    for _ in range(3): m.context_encoder.run_native(ctx, pool=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): m.context_encoder.run_native(ctx, pool=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ops.profile_enable(True)
    for _ in range(3): m.context_encoder.run_native(ctx, pool=True)
    prof = ops.profile_read(); ops.profile_enable(False)
P = B * N
flop = {"conv2": 2*64*128, "conv3": 2*128*256, "conv4": 2*256*512, "conv5": 2*512*1024, "fusion": 2*2048*1024}
out = {k: round(v[0] / 3, 3) for k, v in prof.items()}
tf = {k: round(flop[k] * P / (v[0] / 3 * 1e-3) / 1e12, 1) for k, v in prof.items() if k in flop and v[0] > 0}
print(json.dumps({"prec": prec, "B": B, "N": N, "chunk": chunk, "ms_step": round(ms, 2),
                  "Mpts_s": round(P / ms / 1e3, 1), "stage_ms": out, "stage_tflops": tf}))
