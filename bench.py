#!/usr/bin/env python
"""Benchmark of the LineRefineNet forward hot path (BASELINE.json metric: segments/s & points/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (config.workload): BASELINE.json configs[1] -- encoder forward (shared MLP, fusion, gate,
max+mean pooling -> global_feat) over 4096 segments x 4096 context points per GPU, synthetic
N(0,1) points, random-init weights (PyTorch default init, seed 0), bf16 tensor-core tier.  A "step" is one
pass over that batch.  N > 1 (under torchrun): every rank runs the same per-GPU workload on its
own shard, no data-path collective (segments are independent) -> weak scaling.

  value      : whole-job segments/s with the input batch resident in HBM (device-timed, max over ranks)
  e2e        : same metric through the public API from pinned HOST buffers (pointnet_refine_b200.stream.
               HostEncoderPipeline): H2D copy of the context in 512-segment chunks overlapped with the encoder,
               D2H read of global_feat, every step
  roofline   : fusion GEMM kernel (72.7 % of the FLOPs): algorithmic FLOP per launch / mean launch
               duration from CUDA events recorded around each launch (lrn_profile_*), against the
               sustained bf16 peak of MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the torch-CPU port of the reference (oracle/torch_port.py; the
               reference itself cannot travel to the GPU box) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_POINT_ENCODER = 5_587_584
FLOP_PER_POINT_FUSION_KERNEL = 2 * (1984 + 64) * 1024   # fusion conv + gate layer 2 = 4,194,304
FLOP_PER_POINT_CHAIN_KERNEL = 2 * (64 * 128 + 128 * 256 + 256 * 512 + 512 * 1024)   # conv2..conv5 = 1,392,640
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 200 ms while active."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(segments: int, points: int, steps: int, warmup: int, threads: int | None = None):
    """Time the torch-CPU port of the reference encoder on `threads` host threads (default: all) (segments/s)."""
    import torch
    from oracle import synth, torch_port
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synth.to_torch(synth.make_state_dict(0))
    ctx = torch.from_numpy(synth.make_inputs(segments, points, seed=1234)[0]).transpose(2, 1)
    with torch.no_grad():
        for _ in range(warmup):
            torch_port.encoder_forward(sd, ctx)
        t0 = time.perf_counter()
        for _ in range(steps):
            torch_port.encoder_forward(sd, ctx)
        dt = (time.perf_counter() - t0) / steps
    return segments / dt, dt, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--segments", type=int, default=4096, help="segments per GPU")
    ap.add_argument("--points", type=int, default=4096, help="context points per segment")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--cpu-sample-segments", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = (f"encoder forward + max/mean pooling (global_feat), {args.segments} segments x {args.points} points "
                f"per GPU (BASELINE.json configs[1])")
    config = {"workload": workload, "segments_per_gpu": args.segments, "points_per_segment": args.points,
              "tier": args.precision, "parallelism": f"dp{world} (independent segment shards, no collective)",
              "l2": "input batch (268 MB/GPU at the default size) and per-wave operand buffer (1.24 GB) exceed the 126 MB L2"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        seg = args.cpu_sample_segments
        v, dt, threads = cpu_reference_run(seg, args.points, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
        sample = f"{seg} segments x {args.points} points per step (torch-CPU port of the reference encoder, fp32)"
        print(json.dumps({
            "impl": "reference", "metric": "segments_per_sec", "value": v, "unit": "segments/s", "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 3)), "warmup": max(1, min(args.warmup, 1)), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "points_per_sec": v * args.points, "config": config,
            "cpu_baseline": {"value": v, "unit": "segments/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ native arm (B200)
    import torch
    import torch.distributed as dist

    import pointnet_refine_b200 as prb
    from pointnet_refine_b200 import _lib, ops      # nothing under oracle/ is imported by the native arm

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    torch.manual_seed(0)                            # random-init weights of the reference architecture (PyTorch default
    model = prb.LineRefineNet()                     # init), BatchNorm affine / running statistics randomised so that the
    with torch.no_grad():                           # folding is not the identity
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0.0, 0.1)
                mod.running_mean.normal_(0.0, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
    model = model.to(dev).eval()
    model.precision = args.precision
    enc = model.context_encoder
    B, N = args.segments, args.points
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ctx = torch.randn(B, N, 4, device=dev, generator=gen)

    def step():
        return enc.run_native(ctx, pool=True)["global_feat"]

    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _lib.launch_counter
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        launches = _lib.launch_counter - launches0
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms_per_step = ms / args.steps
        value = world * B / (ms_per_step * 1e-3)

        # ---- per-kernel timing for the roofline (same workload, events around every launch)
        ops.profile_enable(True)
        for _ in range(min(args.steps, 3)):
            step()
        prof = ops.profile_read()
        ops.profile_enable(False)

        # ---- HBM roofline of the stand-alone point loader + first layer (north_star evidence item; the bf16 tier's
        #      default path has this stage inside the fused chain kernel): 16 B in + 256 B out per point
        rows_hbm = min(B * N, 4 * 1024 * 1024)
        folded = enc.folded()
        pts = ctx.reshape(-1, 4)[:rows_hbm]
        def time_embed(tiled):
            for _ in range(2):
                ops.point_embed(folded, pts, tiled=tiled)
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(5):
                ops.point_embed(folded, pts, tiled=tiled)
            h1.record()
            torch.cuda.synchronize()
            return h0.elapsed_time(h1) / 5
        # bf16 tier: the operand matrix of the default path is tiled (contiguous 16 KB blocks); the row-major figure
        # (128-byte pieces at the 4 KB row pitch, the layout of the per-layer / tf32 paths) is reported next to it
        embed_ms_rowmajor = time_embed(False)
        embed_ms = time_embed(True) if args.precision == "bf16" else embed_ms_rowmajor
        embed_bytes = rows_hbm * (16 + 2 * 64 * (2 if args.precision == "bf16" else 4))

        # ---- end to end through the module API from pinned host buffers
        host_ctx = torch.empty(B, N, 4, dtype=torch.float32).pin_memory()
        host_ctx.copy_(ctx)
        host_out = torch.empty(B, 2048, dtype=torch.float32).pin_memory()

        from pointnet_refine_b200.stream import HostEncoderPipeline
        pipe = HostEncoderPipeline(enc, segments_per_chunk=512)

        def e2e_step():   # pinned host context -> (chunked H2D overlapped with the encoder) -> pinned host global_feat
            pipe.global_feat(host_ctx, host_out)

        for _ in range(2):
            e2e_step()
        barrier()
        e2e_steps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e_value = world * B / (e2e_ms / e2e_steps * 1e-3)

        # ---- informational: the whole LineRefineNet forward (encoder + context_proj + 6 decoder layers + heads)
        full_model = None
        if rank == 0:
            fb = min(B, 296)      # 4 x 74 segments x 4096 points = four full encoder waves, 4 attention items per CTA pair
            line = torch.randn(fb, 32, 3, device=dev, generator=gen)
            for _ in range(2):
                model(ctx[:fb], line)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                model(ctx[:fb], line)
            torch.cuda.synchronize()
            fm = (time.perf_counter() - t0) / 3
            full_model = {"segments_per_sec": fb / fm, "ms_per_forward": fm * 1e3, "segments": fb, "points_per_segment": N,
                          "note": "LineRefineNet.forward -> (6,B,32,3); encoder, context_proj, folded-query cross attention "
                                  "(K/V never materialised), query side (linears, 32x32 self attention, add+LayerNorm, heads) on "
                                  "the sm_100a kernels; point_mlp and the K=3 pos_emb layer are stock PyTorch ops"}
            if B * N >= 1024 * 1024:   # inference_whole_scene.py's setting: 1024 context points per line
                ctx_ws = ctx.reshape(-1, 4)[:1024 * 1024].view(1024, 1024, 4)
                line_ws = torch.randn(1024, 32, 3, device=dev, generator=gen)
                for _ in range(2):
                    model(ctx_ws, line_ws)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    model(ctx_ws, line_ws)
                torch.cuda.synchronize()
                fw = (time.perf_counter() - t0) / 3
                full_model["whole_scene_setting"] = {"segments": 1024, "points_per_segment": 1024, "ms_per_forward": fw * 1e3,
                                                     "segments_per_sec": 1024 / fw}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = load_peaks()
    f_ms, f_n = prof["fusion"]
    total_ms = sum(v[0] for v in prof.values())
    points_per_launch = B * N * min(args.steps, 3) / max(f_n, 1)
    achieved = FLOP_PER_POINT_FUSION_KERNEL * points_per_launch / (f_ms / max(f_n, 1) * 1e-3) / 1e12 if f_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    if args.precision == "tf32":
        peak *= 0.5    # no TF32 peak is measured; half the bf16 figure
    roofline = {"bound": "tensor", "kernel": "gemm_pair_kernel<256,EPI_FUSION> (fusion conv + gate + pooling, cta_group::2)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": f"{peak_kind} bf16_tflops_sustained" + (" x 0.5 (tf32)" if args.precision == "tf32" else ""),
                "traffic": None,
                "kernel_ms_per_launch": f_ms / max(f_n, 1), "kernel_share_of_step": f_ms / total_ms if total_ms else None,
                "stage_ms_per_step": {k: v[0] / min(args.steps, 3) for k, v in prof.items()},
                "whole_encoder_tflops": FLOP_PER_POINT_ENCODER * B * N / (ms_per_step * 1e-3) / 1e12}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            roofline["traffic"] = json.load(f).get("fusion_dram_bytes_per_launch")
    except Exception:
        pass
    # second tensor-bound kernel: the fused conv1..conv5 chain (its tensor-core work = conv2..conv5; contains the
    # widest shared-MLP layer, conv5 512 -> 1024).  Only when the chain kernel is the path taken (bf16 tier, defaults).
    c_ms, c_n = prof["conv5"]
    chain_path = args.precision == "bf16" and all(prof[k][1] == 0 for k in ("conv2", "conv3", "conv4")) and c_n > 0
    if chain_path and c_ms > 0:
        c_ach = FLOP_PER_POINT_CHAIN_KERNEL * B * N * min(args.steps, 3) / (c_ms * 1e-3) / 1e12
        roofline["chain_kernel"] = {"kernel": "chain_pair_kernel<true> (conv1 FMA + conv2..conv5 on tcgen05, conv5 with A in TMEM)",
                                    "achieved": c_ach, "peak": peak, "unit": "TFLOP/s", "frac": c_ach / peak,
                                    "kernel_ms_per_launch": c_ms / c_n, "kernel_share_of_step": c_ms / total_ms if total_ms else None,
                                    "algorithmic_flop_per_point": FLOP_PER_POINT_CHAIN_KERNEL}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        seg = args.cpu_sample_segments
        v, dt, threads = cpu_reference_run(seg, N, 3, 1)
        cpu_baseline = {"value": v, "unit": "segments/s", "cores": threads, "kind": "port",
                        "sample": f"{seg} segments x {N} points per step, 1 warm-up + 3 timed "
                                  f"(torch-CPU port of the reference encoder, fp32)", "ms_per_step": dt * 1e3}
        v1, dt1, _ = cpu_reference_run(2, N, 2, 1, threads=1)       # per-core figure
        cpu_baseline["single_thread"] = {"value": v1, "cores": 1, "sample": f"2 segments x {N} points per step, 1 warm-up + 2 timed"}
        torch.set_num_threads(os.cpu_count() or 1)

    print(json.dumps({
        "metric": "segments_per_sec", "value": value, "unit": "segments/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "points_per_sec": value * N,
        "config": config, "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "segments/s", "h2d_bytes_per_step": B * N * 16,
                "d2h_bytes_per_step": B * 2048 * 4, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
        "roofline": roofline,
        "roofline_hbm": {"bound": "hbm", "kernel": "point_embed_kernel (point loading + conv1 + gate layer 1, stand-alone)",
                         "achieved": embed_bytes / (embed_ms * 1e-3) / 1e9, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                         "frac": embed_bytes / (embed_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), "traffic": None,
                         "points": rows_hbm, "algorithmic_bytes_per_point": embed_bytes // rows_hbm,
                         "layout": "tiled operand matrix (default bf16 path)" if args.precision == "bf16" else "row-major operand rows",
                         "row_major_achieved": embed_bytes / (embed_ms_rowmajor * 1e-3) / 1e9,
                         "note": "torch allocation of the operand rows is outside the events' kernel time but inside "
                                 "the loop; the default bf16 path runs this stage inside chain_pair_kernel"},
        "cpu_baseline": cpu_baseline, "full_model": full_model,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
