"""DDP training step (reference train_dist.py:143-189) on the native encoder path: one process per GPU, NCCL.
Checks that DDP's gradient hooks fire through the custom autograd node (gradients are averaged and identical on
all ranks, parameters stay in sync) and prints the step time.  Launch with torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
import pointnet_refine_b200 as prb

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
model = prb.LineRefineNet().to(dev).train()
model.context_encoder.native_training = True            # opt-in: sm_100a train path
ddp = DDP(model, device_ids=[local], find_unused_parameters=True)          # as train_dist.py:147
opt = torch.optim.Adam(ddp.parameters(), lr=1e-3)
g = torch.Generator(device=dev).manual_seed(100 + rank)                    # different shard per rank
ctx = torch.randn(B, N, 4, device=dev, generator=g)
line = torch.randn(B, 32, 3, device=dev, generator=g)
tgt = 0.1 * torch.randn(B, 32, 3, device=dev, generator=g)
times = []
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    out = ddp(ctx, line)
    loss = sum(torch.nn.functional.l1_loss(out[l], tgt) for l in range(6)) / 6
    loss.backward()
    if it == 0:
        worst = 0.0
        for n, p in model.named_parameters():
            assert p.grad is not None, n
            lo, hi = p.grad.clone(), p.grad.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            worst = max(worst, float((hi - lo).abs().max()))
        assert worst == 0.0, f"gradients differ across ranks after the all-reduce: {worst}"
    opt.step()
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum()
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert float(hi - lo) == 0.0, "parameters diverged across ranks"
if rank == 0:
    print(f"ddp ok: world {world}, {B} segments x {N} points per rank, loss {float(loss.detach()):.4f}, "
          f"step {1e3 * min(times[1:]):.1f} ms -> {world * B / min(times[1:]):.0f} segments/s")
dist.destroy_process_group()
