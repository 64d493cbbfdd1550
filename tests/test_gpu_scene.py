"""GPU: scene preprocessing (tube crop, weighted sampling, centroid normalisation) through the C ABI
against the pinned oracle - counts, sampled scene indices and context rows bit for bit."""
import numpy as np
import pytest
import torch

from oracle import scene_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import pointnet_refine_b200  # noqa: F401  (fails loudly if the library is missing)
    return torch.device("cuda:0")


def _check_against_oracle(scene, lines, seg, N, r, decay, seed):
    idx, ctx, cnt = seg.indices.cpu().numpy(), seg.context.cpu().numpy(), seg.counts.cpu().numpy()
    noisy = seg.noisy_line.cpu().numpy()
    for l, raw in enumerate(lines):
        o_ctx, o_noisy, o_center, o_idx, o_count = so.build_segment(scene, raw, N, r, decay, seed, l)
        assert cnt[l] == o_count, (l, cnt[l], o_count)
        np.testing.assert_array_equal(seg.centers[l], o_center)
        np.testing.assert_array_equal(seg.line32[l], so.resample_polyline(raw, 32))      # device resampling == numpy
        np.testing.assert_array_equal(idx[l], o_idx, err_msg=f"line {l}")
        np.testing.assert_array_equal(ctx[l].view(np.uint32), o_ctx.view(np.uint32), err_msg=f"line {l}")
        np.testing.assert_array_equal(noisy[l], o_noisy)


@pytest.mark.parametrize("S,L,N,r", [(200_000, 10, 256, 1.5), (60_000, 5, 1024, 4.0), (50_000, 4, 2048, 0.3), (30_000, 3, 4096, 2.0)])
def test_segments_match_oracle_bit_for_bit(dev, S, L, N, r):
    from pointnet_refine_b200 import scene as sc
    scene, lines = so.synth_scene(S, L, seed=S % 97)
    lines = lines + [lines[0] + np.array([0.0, 900.0, 0.0])]              # a line with an empty tube
    seg = sc.build_segments(torch.from_numpy(scene).to(dev), lines, N, r, 2.0, seed=11)
    _check_against_oracle(scene, lines, seg, N, r, 2.0, 11)
    cnt = seg.counts.cpu().numpy()
    assert cnt[-1] == 0 and (seg.indices[-1] == -1).all()
    if r >= 1.5:
        assert (cnt[:-1] > N).any()                                           # the weighted branch is exercised
    if r == 0.3:
        assert ((cnt[:-1] > 0) & (cnt[:-1] <= N)).any()                       # ... and the with-replacement branch


def test_sorted_and_unsorted_scene_give_identical_segments(dev):
    """The Morton-sorted copy only changes the order in which the crop visits the points."""
    from pointnet_refine_b200 import scene as sc
    scene, lines = so.synth_scene(150_000, 8, seed=21)
    d_scene = torch.from_numpy(scene).to(dev)
    prep = sc.prepare_scene(d_scene)
    assert torch.equal(prep.sorted_points, d_scene[prep.perm.long()])
    a = sc.build_segments(prep, lines, 512, 1.5, 2.0, seed=9)
    b = sc.build_segments(sc.prepare_scene(d_scene, sort=False), lines, 512, 1.5, 2.0, seed=9)
    c = sc.build_segments(d_scene, lines, 512, 1.5, 2.0, seed=9)
    for x in (b, c):
        assert torch.equal(a.indices, x.indices) and torch.equal(a.context, x.context) and torch.equal(a.counts, x.counts)


def test_capacity_overflow_is_retried_and_seed_matters(dev):
    from pointnet_refine_b200 import scene as sc
    scene, lines = so.synth_scene(80_000, 6, seed=5)
    d_scene = torch.from_numpy(scene).to(dev)
    a = sc.build_segments(d_scene, lines, 512, 2.0, 2.0, seed=3, capacity=1000)   # far too small: exact re-run
    b = sc.build_segments(d_scene, lines, 512, 2.0, 2.0, seed=3)
    assert torch.equal(a.indices, b.indices) and torch.equal(a.context, b.context)
    c = sc.build_segments(d_scene, lines, 512, 2.0, 2.0, seed=4)
    assert not torch.equal(a.indices, c.indices)
    _check_against_oracle(scene, lines, a, 512, 2.0, 2.0, 3)


def test_properties_at_scene_scale(dev):
    """2M points, 64 lines: every sample lies inside its tube (float64 distance on the device), no duplicates where the
    tube holds more than N points, counts equal a brute-force float64 count for some lines."""
    from pointnet_refine_b200 import scene as sc
    scene, lines = so.synth_scene(2_000_000, 64, seed=123, extent=400.0)
    d_scene = torch.from_numpy(scene).to(dev)
    N, r = 1024, 1.0
    seg = sc.build_segments(d_scene, lines, N, r, 2.0, seed=99)
    cnt = seg.counts.cpu().numpy()
    for l in (0, 17, 63):
        dense = torch.from_numpy(sc.resample_polyline(lines[l], 200)).to(dev)
        best = torch.full((scene.shape[0],), float("inf"), dtype=torch.float64, device=dev)
        xyz = d_scene[:, :3].double()
        for k in range(200):
            d = xyz - dense[k]
            best = torch.minimum(best, (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        inside = best.sqrt() < r
        assert int(inside.sum()) == cnt[l]
        idx = seg.indices[l]
        assert bool(inside[idx].all())
        if cnt[l] > N:
            assert idx.unique().numel() == N
        rec = d_scene[idx].double()
        rec[:, :3] -= torch.from_numpy(seg.centers[l]).to(dev)
        assert torch.equal(rec.float(), seg.context[l])


def test_refine_scene_runs_lines_as_one_batch(dev):
    import pointnet_refine_b200 as prb
    from pointnet_refine_b200 import scene as sc
    scene, lines = so.synth_scene(100_000, 9, seed=8)
    d_scene = torch.from_numpy(scene).to(dev)
    torch.manual_seed(0)
    m = prb.LineRefineNet().to(dev).eval()
    refined = sc.refine_scene(m, d_scene, lines, 1024, 1.0, 2.0, seed=1)
    assert refined.shape == (9, 32, 3) and np.isfinite(refined).all()
    seg = sc.build_segments(d_scene, lines, 1024, 1.0, 2.0, seed=1)
    with torch.no_grad():
        one = m(seg.context[3:4], seg.noisy_line[3:4])[-1][0].double().cpu().numpy()      # B = 1, like the reference loop
    np.testing.assert_allclose(refined[3], seg.line32[3] + one, atol=5e-2)


def test_device_resampling_is_bit_equal_to_numpy(dev):
    """lrn_scene_resample against the reference formulation (np.interp / linspace / cumsum) on awkward polylines:
    single vertex, repeated vertices, float32-valued vertices, long lines."""
    from pointnet_refine_b200 import scene as sc
    rs = np.random.default_rng(3)
    lines = []
    for i in range(200):
        nv = int(rs.integers(1, 60)) if i % 17 else 1500
        pts = np.cumsum(rs.normal(0, 2, (nv, 3)), axis=0) + rs.normal(0, 200, 3)
        if i % 5 == 0 and nv > 3:
            pts[nv // 2] = pts[nv // 2 - 1]
        if i % 7 == 0:
            pts = pts.astype(np.float32).astype(np.float64)
        if i % 13 == 0 and nv > 1:
            pts[:] = pts[0]                                        # zero total length
        lines.append(pts)
    scene = torch.zeros(64, 4, device=dev)
    seg = sc.build_segments(scene, lines, 16, 0.5, 2.0, seed=0)
    for l, raw in enumerate(lines):
        p32 = so.resample_polyline(raw, 32)
        np.testing.assert_array_equal(seg.line32[l], p32, err_msg=f"line {l} ({len(raw)} vertices)")
        np.testing.assert_array_equal(seg.centers[l], p32.mean(axis=0))
        np.testing.assert_array_equal(seg.noisy_line[l].cpu().numpy(), (p32 - p32.mean(axis=0)).astype(np.float32))


def test_training_batch_matches_dataset_item(dev):
    """build_training_batch = LaneRefineDataset.__getitem__ steps 3-6 for all samples of a scene (src/dataset.py:204-243)."""
    from pointnet_refine_b200 import scene as sc
    scene, lines = so.synth_scene(80_000, 6, seed=12)
    rs = np.random.default_rng(1)
    gts = [l + rs.normal(0, 0.2, l.shape) for l in lines]
    batch = sc.build_training_batch(torch.from_numpy(scene).to(dev), lines, gts, 512, 2.0, 2.0, seed=7)
    for l, (raw, gt) in enumerate(zip(lines, gts)):
        o_ctx, o_noisy, center, _, _ = so.build_segment(scene, raw, 512, 2.0, 2.0, 7, l)
        target = ((so.resample_polyline(gt, 32) - center) - (so.resample_polyline(raw, 32) - center)).astype(np.float32)
        np.testing.assert_array_equal(batch["context"][l].cpu().numpy(), o_ctx)
        np.testing.assert_array_equal(batch["noisy_line"][l].cpu().numpy(), o_noisy)
        np.testing.assert_array_equal(batch["target_offset"][l].cpu().numpy(), target)
