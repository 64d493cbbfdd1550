"""CPU: the preprocessing restatement (oracle/scene_oracle.py) against fixtures generated from the unmodified reference
(oracle/make_scene_golden.py -> tests/golden/scene_small.npz) and the properties of its RNG contract."""
import math
import os

import numpy as np

from oracle import make_scene_golden as msg
from oracle import scene_oracle as so

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene_small.npz"))


def test_resample_crop_and_weights_match_reference_fixture():
    scene, lines = so.synth_scene(**msg.SCENE)
    r, decay = float(G["crop_radius"]), float(G["decay"])
    for l, raw in enumerate(lines):
        p32, p200 = so.resample_polyline(raw, 32), so.resample_polyline(raw, 200)
        np.testing.assert_array_equal(p32, G[f"line{l}_p32"])
        np.testing.assert_array_equal(p200, G[f"line{l}_p200"])
        cand = so.tube_crop(scene, p200, r)
        np.testing.assert_array_equal(cand, G[f"line{l}_crop"])          # same mask as the reference's KDTree query
        w, _ = so.sampling_weights(scene[cand], p32, decay)
        np.testing.assert_allclose(w / w.sum(), G[f"line{l}_p"], rtol=1e-12, atol=0)   # the p of np.random.choice


def test_det_exp_log_accuracy_and_uniform_range():
    x = np.linspace(-40, 3, 20001)
    assert np.max(np.abs(so.det_exp(x) / np.exp(x) - 1)) < 1e-15
    u = so.det_uniform(7, 3, np.arange(200000))
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3
    assert np.max(np.abs(so.det_log(u) - np.log(u)) / np.maximum(1e-3, np.abs(np.log(u)))) < 1e-15
    assert np.max(np.abs(so.det_log(np.array([1.0, 0.5, 2.0 ** -53])) - np.log([1.0, 0.5, 2.0 ** -53]))) < 1e-14
    # different lines / seeds decorrelate
    assert abs(np.corrcoef(so.det_uniform(7, 3, np.arange(5000)), so.det_uniform(7, 4, np.arange(5000)))[0, 1]) < 0.05
    assert abs(np.corrcoef(so.det_uniform(7, 3, np.arange(5000)), so.det_uniform(8, 3, np.arange(5000)))[0, 1]) < 0.05


def test_sampling_distribution_matches_reference_draws():
    """Inclusion frequencies of the Efraimidis-Spirakis sampler over 4000 seeds vs 4000 draws of the reference's
    np.random.choice(replace=False, p) on the same 60 candidates (choose 20)."""
    scene, lines = so.synth_scene(**msg.SCENE)
    p32 = so.resample_polyline(lines[0], 32)
    cand = G["line0_crop"][:60]
    trials = int(G["dist_trials"])
    counts = np.zeros(60, np.int64)
    pos = {int(c): i for i, c in enumerate(cand)}
    for seed in range(trials):
        idx = so.sample_indices(scene, cand, p32, 20, float(G["decay"]), seed, 0)
        assert len(set(idx.tolist())) == 20                       # without replacement
        counts[[pos[int(i)] for i in idx]] += 1
    f_ref, f_new = G["dist_counts"] / trials, counts / trials
    sigma = np.sqrt(np.maximum(f_ref * (1 - f_ref), 1e-4) * 2 / trials)
    assert np.all(np.abs(f_ref - f_new) < 4.5 * sigma), (np.abs(f_ref - f_new) / sigma).max()
    assert abs(counts.sum() - 20 * trials) == 0


def test_small_and_empty_crops_follow_reference_rules():
    scene, lines = so.synth_scene(2000, 1, 9)
    far = lines[0] + np.array([0.0, 500.0, 0.0])
    ctx, noisy, center, idx, count = so.build_segment(scene, far, 64, 1.0, 2.0, 1, 0)
    assert count == 0 and (idx == -1).all()
    np.testing.assert_allclose(ctx[:, :3], np.broadcast_to(-center, (64, 3)).astype(np.float32))     # zeros - center
    assert (ctx[:, 3] == 0).all()
    ctx, noisy, center, idx, count = so.build_segment(scene, lines[0], 4096, 0.3, 2.0, 1, 0)
    assert 0 < count <= 4096 and set(idx.tolist()) <= set(so.tube_crop(scene, so.resample_polyline(lines[0], 200), 0.3).tolist())
    np.testing.assert_allclose(noisy.astype(np.float64).mean(axis=0), 0, atol=1e-5)
    # reproducible and seed / line dependent
    a = so.build_segment(scene, lines[0], 128, 2.0, 2.0, 5, 0)[3]
    assert np.array_equal(a, so.build_segment(scene, lines[0], 128, 2.0, 2.0, 5, 0)[3])
    assert not np.array_equal(a, so.build_segment(scene, lines[0], 128, 2.0, 2.0, 6, 0)[3])
    assert not np.array_equal(a, so.build_segment(scene, lines[0], 128, 2.0, 2.0, 5, 1)[3])
