"""Scene preprocessing: device time of build_segments for a whole scene vs the per-line host
algorithm (oracle/scene_oracle.py, and the reference's KD-tree formulation when scipy is importable), and
refine_scene (preprocessing + one batched forward) vs a B = 1 loop like inference_whole_scene.py:124-142."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pointnet_refine_b200 as prb
from pointnet_refine_b200 import scene as sc
from oracle import scene_oracle as so
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
r = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
dev = torch.device("cuda:0")
scene, lines = so.synth_scene(S, L, seed=1, extent=300.0)
d_scene = torch.from_numpy(scene).to(dev)
def dev_time(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
t_prep = dev_time(lambda: sc.prepare_scene(d_scene))
prep = sc.prepare_scene(d_scene)
t_unsorted = dev_time(lambda: sc.build_segments(sc.prepare_scene(d_scene, sort=False), lines, N, r, 2.0, seed=1))
t_build = dev_time(lambda: sc.build_segments(prep, lines, N, r, 2.0, seed=1))
t0 = time.perf_counter()
for raw in lines: sc.resample_polyline(raw, 32); sc.resample_polyline(raw, 200)
t_resample = time.perf_counter() - t0
seg = sc.build_segments(d_scene, lines, N, r, 2.0, seed=1)
cand = int(seg.counts.sum())
# host baselines on a few lines
k = min(L, 4)
t0 = time.perf_counter()
for l in range(k): so.build_segment(scene, lines[l], N, r, 2.0, 1, l)
t_oracle = (time.perf_counter() - t0) / k
t_kd = None
try:
    from scipy.spatial import KDTree
    t0 = time.perf_counter()
    for l in range(k):   # the reference's formulation: KD-tree on the dense line, query every scene point (src/dataset.py:217-222)
        d, _ = KDTree(so.resample_polyline(lines[l], 200)).query(scene[:, :3])
        c = scene[d < r]
        if len(c) > N:
            KDTree(so.resample_polyline(lines[l], 32)).query(c[:, :3])
    t_kd = (time.perf_counter() - t0) / k
except ImportError:
    pass
torch.manual_seed(0)
m = prb.LineRefineNet().to(dev).eval()
t_refine = dev_time(lambda: sc.refine_scene(m, prep, lines, N, r, 2.0, seed=1), n=2)
with torch.no_grad():
    def loop():
        for l in range(min(L, 32)):
            m(seg.context[l:l + 1], seg.noisy_line[l:l + 1])[-1].cpu()
    t_loop = dev_time(loop, n=2) / min(L, 32)
print(json.dumps({"scene_points": S, "lines": L, "N": N, "crop_radius": r, "candidates": cand,
                  "prepare_scene_ms": round(t_prep * 1e3, 2), "build_segments_unsorted_ms": round(t_unsorted * 1e3, 2), "build_segments_ms": round(t_build * 1e3, 2), "numpy_resampling_of_all_lines_ms": round(t_resample * 1e3, 2), "ms_per_line": round(t_build * 1e3 / L, 4),
                  "scene_points_x_lines_per_s": round(S * L / t_build, 1),
                  "host_oracle_ms_per_line": round(t_oracle * 1e3, 1), "host_kdtree_ms_per_line": None if t_kd is None else round(t_kd * 1e3, 1),
                  "refine_scene_ms": round(t_refine * 1e3, 2), "b1_forward_loop_ms_per_line": round(t_loop * 1e3, 3),
                  "refine_scene_lines_per_s": round(L / t_refine, 1)}))
