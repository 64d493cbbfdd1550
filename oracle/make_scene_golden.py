"""Generate tests/golden/scene_small.npz from the UNMODIFIED reference preprocessing (src/dataset.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Build container only (``python -m oracle.make_scene_golden``).
For a deterministic synthetic scene (oracle/scene_oracle.synth_scene - regenerated from the seed by the tests) it stores,
per line: the reference's resample_polyline outputs (32 and 200 points), the reference's crop mask (KDTree query) and
the probability vector the reference's weighted_sampling hands to np.random.choice (captured by wrapping it - the draw
itself uses the global unseeded stream and cannot be pinned).  Plus, for one small candidate set, the inclusion counts
of 4000 reference draws: the distribution the RNG contract of oracle/scene_oracle.py has to reproduce.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scene_small.npz")
SCENE = dict(num_points=20000, num_lines=3, seed=42)
CROP_RADIUS, DECAY, N_SAMPLES = 1.5, 2.0, 256


def main():
    sys.path.insert(0, REF)
    from scipy.spatial import KDTree
    from src import dataset as ref
    from oracle import scene_oracle as so

    scene, lines = so.synth_scene(**SCENE)
    out = {"crop_radius": CROP_RADIUS, "decay": DECAY, "n_samples": N_SAMPLES}
    real_choice = np.random.choice
    for l, raw in enumerate(lines):
        p32, p200 = ref.resample_polyline(raw, 32), ref.resample_polyline(raw, 200)
        d, _ = KDTree(p200).query(scene[:, :3])
        mask = d < CROP_RADIUS
        cand = scene[mask]
        captured = {}

        def spy(a, size=None, replace=True, p=None):
            captured["p"] = None if p is None else np.array(p)
            return real_choice(a, size, replace=replace, p=p)
        np.random.choice = spy
        try:
            sampled = ref.weighted_sampling(cand, p32, num_samples=N_SAMPLES, decay_scale=DECAY)
        finally:
            np.random.choice = real_choice
        assert sampled.shape == (N_SAMPLES, 4) and captured["p"] is not None
        out[f"line{l}_p32"], out[f"line{l}_p200"] = p32, p200
        out[f"line{l}_crop"] = np.nonzero(mask)[0].astype(np.int64)
        out[f"line{l}_p"] = captured["p"]
    # distribution of the reference's draw on a small set: 60 candidates nearest to line 0, choose 20, 4000 trials
    p32 = out["line0_p32"]
    cand_idx = out["line0_crop"][:60]
    cand = scene[cand_idx]
    counts = np.zeros(60, np.int64)
    np.random.seed(12345)
    probs = {}

    def spy2(a, size=None, replace=True, p=None):
        probs["p"] = np.array(p)
        r = real_choice(a, size, replace=replace, p=p)
        counts[r] += 1
        return r
    np.random.choice = spy2
    try:
        for _ in range(4000):
            ref.weighted_sampling(cand, p32, num_samples=20, decay_scale=DECAY)
    finally:
        np.random.choice = real_choice
    out["dist_counts"], out["dist_trials"], out["dist_p"] = counts, 4000, probs["p"]
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", {k: getattr(v, "shape", v) for k, v in out.items() if "crop" in k})


if __name__ == "__main__":
    main()
