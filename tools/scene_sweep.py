"""BASELINE.json configs[2]: inference_whole_scene-style sweep, S segments x 2048 points, data-parallel over the ranks
(contiguous segment shards, no collective; `python tools/scene_sweep.py` or under torchrun).  Each rank streams its
shard through the whole LineRefineNet forward (or the encoder alone, --encoder) in chunks; the synthetic chunk is
generated on the device (the full input would be 32.8 GB of fp32).  Device time, barrier on both sides, max over ranks.

  python tools/scene_sweep.py [--segments 1000000] [--points 2048] [--chunk 512] [--encoder]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import pointnet_refine_b200 as prb
from pointnet_refine_b200.shard import segment_shard

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=1_000_000)
ap.add_argument("--points", type=int, default=2048)
ap.add_argument("--chunk", type=int, default=1184, help="segments per forward call (1184 x 2048 points = two decoder passes of four full encoder waves)")
ap.add_argument("--encoder", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay the chunk forward as a CUDA graph")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = prb.LineRefineNet().to(dev).eval()
mine = segment_shard(args.segments, rank, world)
gen = torch.Generator(device=dev).manual_seed(1234 + rank)
ctx = torch.randn(args.chunk, args.points, 4, device=dev, generator=gen)
line = torch.randn(args.chunk, 32, 3, device=dev, generator=gen)
acc = torch.zeros((), device=dev, dtype=torch.float64)
graphed = prb.GraphedLineRefineNet(m, args.chunk, args.points) if args.graph and not args.encoder else None

def run(n_seg):
    done = 0
    while done < n_seg:
        n = min(args.chunk, n_seg - done)
        if args.encoder:
            out = m.context_encoder.run_native(ctx[:n], pool=True)["global_feat"]
        elif graphed is not None and n == args.chunk:
            out = graphed(ctx, line)[-1]
        else:
            out = m(ctx[:n], line[:n])[-1]
        acc.add_(out.double().sum())            # consume the result on the device
        done += n

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

with torch.no_grad():
    run(4 * args.chunk)                           # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(len(mine))
    e1.record()
    barrier()
t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "scene sweep (BASELINE configs[2])", "segments": args.segments, "points": args.points, "n_gpus": world,
                      "what": "encoder + pooling" if args.encoder else "LineRefineNet.forward (6,B,32,3)", "chunk": args.chunk, "cuda_graph": bool(graphed),
                      "seconds": round(float(t), 3), "segments_per_sec": round(args.segments / float(t), 1),
                      "points_per_sec": round(args.segments * args.points / float(t), 1), "finite": bool(torch.isfinite(acc))}))
if world > 1:
    dist.destroy_process_group()
