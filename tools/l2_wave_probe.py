"""Wave size vs throughput / DRAM traffic of the two-kernel encoder (VERDICT r01 item 3 experiment): with waves small
enough for the 4 KB/point operand rows to stay in the 126 MB L2 the fusion kernel reads them from L2, not HBM - at the price
of one tile per CTA pair per launch (no cross-tile pipelining, launch gaps).  usage: l2_wave_probe.py [segments points]"""
import sys, os, json, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnet_refine_b200 as prb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
m = prb.LineRefineNet().to(dev).eval()
enc = m.context_encoder
ctx = torch.randn(B, N, 4, device=dev)
def clocks():
    out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
    return out.strip()
res = []
with torch.no_grad():
    ref = None
    for rows in (18944, 37888, 75776, 151552, 303104):
        enc.chunk_rows = rows
        for _ in range(2): gf = enc.run_native(ctx, pool=True)["global_feat"]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        samples = []
        stop = False
        def samp():
            while not stop:
                samples.append(clocks()); time.sleep(0.1)
        th = threading.Thread(target=samp); th.start()
        e0.record()
        reps = 20
        for _ in range(reps): gf = enc.run_native(ctx, pool=True)["global_feat"]
        e1.record(); torch.cuda.synchronize()
        stop = True; th.join()
        ms = e0.elapsed_time(e1) / reps
        if ref is None: ref = gf.clone()
        res.append({"wave_points": rows, "operand_bytes_per_wave": rows * 4096, "ms": ms, "segments_per_sec": B / ms * 1e3,
                    "max_abs_vs_first": float((gf - ref).abs().max()), "clock_power_samples": samples[1:-1][:6]})
        print(json.dumps(res[-1]), flush=True)
