"""Upper bound for keeping the operand rows on the chip: the fusion kernel ALONE, fully pipelined
(16 tiles per CTA pair per launch), reading its 4 KB/point operand rows (a) from HBM as in the product path and (b) from an
L2-resident ring of R row tiles (every launch re-reads the same R x 256 rows: R = 74 -> 77.6 MB, R = 37 -> 38.8 MB), with
clocks and board power sampled under load.  Needs the -DLRN_TIMELINE tuning build (this script compiles it and re-executes
itself with LRN_B200_LIB pointing at it); the ring makes the RESULTS meaningless - only time, clocks and power are read.
python tools/l2_ring_probe.py [segments points]"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "LRN_B200_LIB" not in os.environ:
    sys.path.insert(0, os.path.join(ROOT, "pointnet_refine_b200"))
    import build as lrn_build
    lib = lrn_build.build(timeline=True)
    for ring, skip in (("0", "0"), ("0", "1"), ("37", "1"), ("74", "1"), ("0", "1"), ("37", "1")):
        env = dict(os.environ, LRN_B200_LIB=lib, LRN_DBG_RING=ring, LRN_DBG_SKIP_CHAIN=skip, LRN_DBG_LAYER="9")
        subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, check=True)
    sys.exit(0)

import torch  # noqa: E402
import pointnet_refine_b200 as prb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = prb.LineRefineNet().to(dev).eval()
enc = m.context_encoder
ctx = torch.randn(B, N, 4, device=dev)


def sample():
    out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                         capture_output=True, text=True).stdout.strip().split(",")
    return float(out[0]), float(out[1])


with torch.no_grad():
    skip = os.environ["LRN_DBG_SKIP_CHAIN"]
    os.environ["LRN_DBG_SKIP_CHAIN"] = "0"      # getenv is read per launch: fill the operand buffer with real activations first
    enc.run_native(ctx, pool=True)
    torch.cuda.synchronize()
    os.environ["LRN_DBG_SKIP_CHAIN"] = skip
    for _ in range(3):
        enc.run_native(ctx, pool=True)
    torch.cuda.synchronize()
    samples, stop = [], False

    def samp():
        while not stop:
            samples.append(sample())
            time.sleep(0.15)
    th = threading.Thread(target=samp)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ.get("LRN_PROBE_REPS", "200"))      # ~6 s of load: long enough for the power cap to settle
    e0.record()
    for _ in range(reps):
        enc.run_native(ctx, pool=True)
    e1.record()
    torch.cuda.synchronize()
    stop = True
    th.join()
ms = e0.elapsed_time(e1) / reps
mid = samples[len(samples) // 4:] or samples
flop = 4194304.0 * B * N
print(json.dumps({"ring_tiles": int(os.environ["LRN_DBG_RING"]), "chain_skipped": skip == "1", "segments": B, "points": N,
                  "ms_per_pass": ms, "fusion_tflops_if_alone": flop / ms / 1e9 if skip == "1" else None,
                  "segments_per_sec": B / ms * 1e3, "sm_mhz_median": sorted(s[0] for s in mid)[len(mid) // 2],
                  "power_w_median": sorted(s[1] for s in mid)[len(mid) // 2], "samples": len(samples)}), flush=True)
