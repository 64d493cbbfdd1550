"""Train step with the native loop machinery (FlatAdam + fused deep-supervision loss) vs torch.optim.Adam + the loss loop."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnet_refine_b200 as prb
from pointnet_refine_b200.optim import FlatAdam, deep_supervision_l1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
res = {}
for flat in (True, False):
    torch.manual_seed(0)
    m = prb.LineRefineNet().to(dev).train()
    m.context_encoder.native_training = True
    opt = FlatAdam(m.parameters(), lr=1e-3) if flat else torch.optim.Adam(m.parameters(), lr=1e-3)
    ctx = torch.randn(B, N, 4, device=dev); line = torch.randn(B, 32, 3, device=dev); tgt = torch.randn(B, 32, 3, device=dev)
    def step():
        opt.zero_grad()
        out = m(ctx, line)
        loss = deep_supervision_l1(out, tgt) if flat else sum(torch.nn.functional.l1_loss(out[l], tgt) for l in range(6)) / 6
        loss.backward(); opt.step()
        return loss
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): l = step()
    torch.cuda.synchronize()
    res["flat_adam+fused_l1" if flat else "torch_adam+loss_loop"] = {"ms_per_step": round((time.perf_counter() - t0) / 4 * 1e3, 1), "loss": float(l.detach())}
    del m, opt; torch.cuda.empty_cache()
print(json.dumps({"B": B, "N": N, **res}))
