"""Eager vs CUDA-graph replay of the whole training step (GraphedTrainStep) at B x N points; optional 2+ ranks under
torchrun (FlatDataParallel inside the graph).  python tools/graph_train_bench.py [B N]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pointnet_refine_b200 as prb  # noqa: E402
from pointnet_refine_b200.optim import FlatAdam, deep_supervision_l1  # noqa: E402

B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 1024)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = prb.LineRefineNet().to(dev).train()
m.context_encoder.native_training = True
net = prb.FlatDataParallel(m) if world > 1 else m
opt = FlatAdam(m.parameters(), lr=1e-3, capturable=True)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ctx = torch.randn(B, N, 4, device=dev, generator=g)
line = torch.randn(B, 32, 3, device=dev, generator=g)
tgt = 0.1 * torch.randn(B, 32, 3, device=dev, generator=g)


def eager():
    opt.zero_grad()
    loss = deep_supervision_l1(net(ctx, line), tgt)
    loss.backward()
    opt.step()
    return loss.detach()     # keep no autograd graph alive: its AccumulateGrad nodes would pin the default stream


def timed(fn, steps=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        out = fn()
    host = (time.perf_counter() - t0) / steps * 1e3
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, host, out


res = {"B": B, "N": N, "world": world, "rank": rank}
ms, host, loss = timed(eager)
res["eager_ms"], res["eager_host_enqueue_ms"], res["eager_loss"] = ms, host, float(loss)
t0 = time.perf_counter()
step = prb.GraphedTrainStep(net, opt, ctx, line, tgt)
res["capture_s"] = time.perf_counter() - t0
ms, host, out = timed(lambda: step(ctx, line, tgt))
res["graph_ms"], res["graph_host_enqueue_ms"], res["graph_loss"] = ms, host, float(out[0])
res["steps_taken"] = opt.steps_taken
res["peak_mem_gb"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
print(json.dumps(res), flush=True)
del out
step.close()
if world > 1:
    dist.destroy_process_group()
