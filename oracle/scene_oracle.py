"""CPU restatement of the reference's per-line preprocessing: tube crop, weighted
sampling without replacement, centroid normalisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: src/dataset.py:8-30 (resample_polyline), :78-130 (weighted_sampling), :214-237 (crop + normalise in
LaneRefineDataset.__getitem__), inference_whole_scene.py:95-121 (process_single_line, same logic).

What is pinned against the reference (oracle/make_scene_golden.py -> tests/golden/scene_small.npz): the resampled
polylines, the crop mask (the reference's KDTree query vs the brute-force minimum here) and the sampling
probabilities `p` the reference hands to np.random.choice (captured by wrapping np.random.choice).  What cannot be
pinned: the draw itself - the reference uses the global, unseeded np.random stream.  The RNG CONTRACT of this
framework replaces it (same distribution, reproducible, order-independent, so that a GPU can evaluate it per point):

  u(seed, line, i) = ((splitmix64(seed + GOLD*(line+1) + MIX*(i+1)) >> 11) + 0.5) * 2^-53            in (0, 1)
  more than N candidates: Efraimidis-Spirakis keys  k_i = det_log(u_i) / w_i  (w_i > 0, unnormalised weights);
      the N largest keys are the sample, in descending key order (ties: smaller scene index first).  Sampling
      without replacement with probabilities proportional to w - the distribution of np.random.choice(replace=False, p=w/sum w).
  1..N candidates: draw j takes candidate floor(u(seed, line, 2^40 + j) * count) of the candidates in ascending scene
      index order (uniform with replacement, src/dataset.py:89-91).
  no candidates: N zero points (src/dataset.py:87-88).

det_exp / det_log are fixed sequences of IEEE-754 double operations (no fused multiply-add, no libm), so that numpy and
CUDA produce the same bits and the selected indices are bit-exact by construction.
"""
from __future__ import annotations

import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
MIX = np.uint64(0xD1B54A32D192ED03)
REPLACE_BASE = 1 << 40
LN2_HI = 6.93147180369123816490e-01
LN2_LO = 1.90821492927058770002e-10
INV_LN2 = 1.44269504088896338700e+00


# ------------------------------------------------------------------------------------------------ polylines
def resample_polyline(points, num_points=32):
    """Arc-length-uniform linear resampling to `num_points` (src/dataset.py:8-30); fewer than two points -> zeros."""
    points = np.asarray(points, np.float64)
    if len(points) < 2:
        return np.zeros((num_points, 3))
    seg = np.linalg.norm(points[1:] - points[:-1], axis=1)
    cum = np.concatenate(([0.0], np.cumsum(seg)))
    t = np.linspace(0, cum[-1], num_points)
    return np.stack([np.interp(t, cum, points[:, k]) for k in range(3)], axis=1)


# ------------------------------------------------------------------------------------------------ RNG contract
def splitmix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        z = x + GOLD
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def det_uniform(seed, line, idx):
    idx = np.asarray(idx, np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) + GOLD * np.uint64(line + 1) + MIX * (idx + np.uint64(1))
    return ((splitmix64(x) >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53


def det_exp(x):
    """exp(x) for |x| < 700 as plain double operations: x = n ln2 + r, degree-13 Taylor polynomial of r (Horner)."""
    x = np.asarray(x, np.float64)
    n = np.rint(x * INV_LN2)
    r = (x - n * LN2_HI) - n * LN2_LO
    p = np.full_like(r, 1.0 / 6227020800.0)
    for c in (1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
              1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        p = p * r + c
    return np.ldexp(p, n.astype(np.int32))


def det_log(u):
    """log(u) for u in (0, 1]: u = m 2^e with m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1), log m = 2 s (1 + s^2/3 + ... + s^24/25)."""
    u = np.asarray(u, np.float64)
    m, e = np.frexp(u)                       # m in [0.5, 1)
    low = m < 0.70710678118654752440
    m = np.where(low, m * 2.0, m)
    e = np.where(low, e - 1, e).astype(np.float64)
    s = (m - 1.0) / (m + 1.0)
    s2 = s * s
    p = np.full_like(s, 1.0 / 25.0)
    for k in (23, 21, 19, 17, 15, 13, 11, 9, 7, 5, 3, 1):
        p = p * s2 + 1.0 / k
    return (e * LN2_HI + (2.0 * s) * p) + e * LN2_LO


# ------------------------------------------------------------------------------------------------ crop + weights
def min_distance(xyz, line_points, chunk=65536):
    """Euclidean distance (float64) of every row of xyz (S,3) to its nearest row of line_points (M,3): what
    KDTree(line_points).query(xyz)[0] returns (src/dataset.py:217-219), as sqrt((dx^2 + dy^2) + dz^2)."""
    xyz = np.asarray(xyz, np.float64)
    lp = np.asarray(line_points, np.float64)
    out = np.empty(len(xyz))
    for s in range(0, len(xyz), chunk):
        d = xyz[s:s + chunk, None, :] - lp[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        out[s:s + chunk] = np.sqrt(d2.min(axis=1))
    return out


def tube_crop(scene, dense_line, crop_radius):
    """Indices (ascending) of the scene points closer than crop_radius to the 200-point polyline (src/dataset.py:214-222)."""
    return np.nonzero(min_distance(scene[:, :3], dense_line) < crop_radius)[0]


def sampling_weights(cand, noisy32, decay_scale):
    """Unnormalised weights of src/dataset.py:94-112: exp(-d/decay) (float64, d = distance to the 32-point line) times
    0.5 + normalised intensity (float32 arithmetic, like the reference's float32 point array)."""
    d = min_distance(cand[:, :3], noisy32)
    inten = cand[:, 3].astype(np.float32)
    lo, hi = inten.min(), inten.max()
    if hi > lo:
        norm = (inten - lo) / ((hi - lo) + np.float32(1e-6))
    else:
        norm = np.full_like(inten, 0.5)
    iw = np.float32(0.5) + norm
    return det_exp(-d / decay_scale) * iw.astype(np.float64), d


def sample_indices(scene, cand_idx, noisy32, num_samples, decay_scale, seed, line):
    """Scene indices (num_samples,) int64 under the RNG contract in the module docstring; -1 for the zero points of an
    empty crop."""
    count = len(cand_idx)
    if count == 0:
        return np.full(num_samples, -1, np.int64)
    if count <= num_samples:
        u = det_uniform(seed, line, REPLACE_BASE + np.arange(num_samples))
        pick = np.minimum((u * count).astype(np.int64), count - 1)
        return cand_idx[pick].astype(np.int64)
    w, _ = sampling_weights(scene[cand_idx], noisy32, decay_scale)
    if not (w.sum() >= 1e-6):       # the reference falls back to uniform weights (src/dataset.py:115-117)
        w = np.ones_like(w)
    keys = det_log(det_uniform(seed, line, cand_idx)) / w
    order = np.lexsort((cand_idx, -keys))                 # key descending, then scene index ascending
    return cand_idx[order[:num_samples]].astype(np.int64)


def build_segment(scene, raw_line, num_samples, crop_radius, decay_scale, seed, line):
    """One line of the scene -> (context (N,4) f32, noisy_centered (32,3) f32, center (3,) f64, indices (N,) i64, count)
    following LaneRefineDataset.__getitem__ steps 3-6 / process_single_line steps 1-4."""
    scene = np.asarray(scene, np.float32)
    noisy32 = resample_polyline(raw_line, 32)
    dense = resample_polyline(raw_line, 200)
    cand = tube_crop(scene, dense, crop_radius) if len(scene) else np.zeros(0, np.int64)
    idx = sample_indices(scene, cand, noisy32, num_samples, decay_scale, seed, line)
    center = noisy32.mean(axis=0)
    pts = np.where(idx[:, None] >= 0, scene[np.maximum(idx, 0)].astype(np.float64), 0.0)
    context = np.concatenate([pts[:, :3] - center, pts[:, 3:4]], axis=1).astype(np.float32)
    return context, (noisy32 - center).astype(np.float32), center, idx, len(cand)


def synth_scene(num_points, num_lines, seed, extent=60.0):
    """Deterministic synthetic scene: ground points + dense stripes of bright points along `num_lines` wavy lane lines,
    and one noisy raw polyline (8-14 vertices, not uniformly spaced) per lane."""
    rs = np.random.Generator(np.random.PCG64(seed))
    lines = []
    n_lane = num_points // 3
    lane_pts = []
    for l in range(num_lines):
        y0 = rs.uniform(-extent / 3, extent / 3)
        amp, ph = rs.uniform(0.5, 3.0), rs.uniform(0, 6.28)
        f = lambda x: y0 + amp * np.sin(x / 15.0 + ph)
        nv = int(rs.integers(8, 15))
        xs = np.sort(rs.uniform(-extent / 2, extent / 2, nv))
        raw = np.stack([xs, f(xs) + rs.normal(0, 0.3, nv), rs.normal(0, 0.05, nv)], axis=1)
        lines.append(raw)
        k = n_lane // num_lines
        x = rs.uniform(-extent / 2, extent / 2, k)
        lane_pts.append(np.stack([x, f(x) + rs.normal(0, 0.08, k), rs.normal(0, 0.03, k), np.clip(np.round(rs.gamma(6, 12, k)), 0, 255)], axis=1))
    g = num_points - sum(len(p) for p in lane_pts)
    ground = np.stack([rs.uniform(-extent / 2, extent / 2, g), rs.uniform(-extent / 2, extent / 2, g), rs.normal(0, 0.1, g),
                       np.clip(np.round(rs.gamma(2, 8, g)), 0, 255)], axis=1)
    scene = np.concatenate(lane_pts + [ground]).astype(np.float32)
    scene = scene[rs.permutation(len(scene))]
    return scene, lines
