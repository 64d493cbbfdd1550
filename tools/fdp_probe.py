"""Diagnostic (torchrun): train-step time under FlatDataParallel after a large inference leg; profiler view of one step."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import pointnet_refine_b200 as prb
from pointnet_refine_b200 import optim as lrn_optim
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
big = len(sys.argv) > 1 and sys.argv[1] == "big"
if big:
    mi = prb.LineRefineNet().to(dev).eval()
    c = torch.randn(4096, 4096, 4, device=dev); l = torch.randn(4096, 32, 3, device=dev)
    with torch.no_grad():
        for _ in range(3): mi.context_encoder.run_native(c, pool=True)
        mi(c, l)
    torch.cuda.synchronize(); del c, l
    torch.cuda.empty_cache()
g = torch.Generator(device=dev).manual_seed(100 + rank)
B, N = 1024, 1024
ctx = torch.randn(B, N, 4, device=dev, generator=g); line = torch.randn(B, 32, 3, device=dev, generator=g); tgt = 0.1 * torch.randn(B, 32, 3, device=dev, generator=g)
def run(tag, **kw):
    torch.manual_seed(0)
    m = prb.LineRefineNet().to(dev).train(); m.context_encoder.native_training = True
    net = prb.FlatDataParallel(m, **kw) if tag != "single" else m
    opt = lrn_optim.FlatAdam(m.parameters(), lr=1e-3)
    def one():
        opt.zero_grad(); loss = lrn_optim.deep_supervision_l1(net(ctx, line), tgt); loss.backward(); opt.step()
    for _ in range(4): one()
    torch.cuda.synchronize(); dist.barrier(device_ids=[local]); t0 = time.perf_counter()
    for it in range(8): one()
    torch.cuda.synchronize(); t8 = (time.perf_counter() - t0) / 8
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        one(); torch.cuda.synchronize()
    ev = prof.key_averages()
    tot = sum(e.device_time_total for e in ev)
    nccl = sum(e.device_time_total for e in ev if "nccl" in e.key.lower())
    print(f"rank {rank} {tag} {kw} big={big}: {1e3 * t8:.1f} ms/step; profiled step: device busy {tot / 1e3:.1f} ms, of which nccl {nccl / 1e3:.2f} ms", flush=True)
run("single")
run("flat", overlap=True)
run("flat", overlap=False)
dist.destroy_process_group()
