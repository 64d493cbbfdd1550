"""cuBLAS 8192^3 matmul throughput, bf16 and tf32, burst (best of 10) and sustained (back to back for ~3 s), with the SM
clock nvidia-smi reports while the sustained loop runs.  Same method as MEASURED_PEAKS.json; gives the tf32 tier a
measured denominator and the bf16 figure of THIS box next to tools/mma_probe's MMA-only stream."""
import json, subprocess, sys, threading, time
import torch

dev = torch.device("cuda:0")
n = 8192
out = {}
for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dt)
    b = torch.randn(n, n, device=dev, dtype=dt)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    lines = []
    proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                            stdout=subprocess.PIPE, text=True)
    threading.Thread(target=lambda: lines.extend(proc.stdout), daemon=True).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        iters += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    proc.terminate()
    sus = 2 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    clk = sorted(float(l.split(",")[0]) for l in lines[5:] if "," in l)
    pw = sorted(float(l.split(",")[1]) for l in lines[5:] if "," in l)
    out[name] = {"burst_tflops": best, "sustained_tflops": sus, "sm_mhz_median": clk[len(clk) // 2] if clk else None,
                 "power_w_median": pw[len(pw) // 2] if pw else None}
    del a, b
print(json.dumps(out))
