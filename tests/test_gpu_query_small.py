"""Query side of the decoder for few polyline rows (the whole-scene loop's B = 1,
reference inference_whole_scene.py:130-139): the fp32 rows_linear kernel and its fused producers, the split merge of the
attention kernel, point_mlp / pos_emb through them, and the claim that such a forward launches only this library's
kernels (no cuBLAS / cutlass / fmha / ATen elementwise kernel)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import lrn_oracle as orc  # noqa: E402
from oracle import synth  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _rna_tf32(t):
    """fp32 tensor -> nearest TF32 value (ties away from zero), the rounding of cvt.rna.tf32.f32"""
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K", [(1, 8, 32), (32, 256, 256), (33, 512, 256), (224, 2048, 256), (32, 256, 2048), (96, 256, 1024),
                                   (32, 128, 64), (65, 256, 128)])
def test_rows_linear_matches_fp64(dev, M, N, K):
    from pointnet_refine_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M * 7 + N + K)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, x2, w, b = r(M, K), r(M, K), r(N, K) / K ** 0.5, r(N)
    ref = x.double() @ w.double().T + b.double()
    tol = 2e-5 * max(1.0, float(ref.abs().max()))
    assert float((ops.rows_linear(x, w, b).double() - ref).abs().max()) <= tol
    assert float((ops.rows_linear(x, w, None).double() - (ref - b.double())).abs().max()) <= tol
    assert float((ops.rows_linear(x, w, b, relu=True).double() - ref.clamp_min(0)).abs().max()) <= tol
    ref2 = (x.double() + x2.double()) @ w.double().T + b.double()
    assert float((ops.rows_linear(x, w, b, add=x2).double() - ref2).abs().max()) <= 2 * tol
    o16 = ops.rows_linear(x, w, b, out_dtype=torch.bfloat16)
    assert o16.dtype == torch.bfloat16 and float((o16.double() - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max()) + 1e-5
    # the K = 3 first layer of an MLP fused into the operand load (pos_emb, point_mlp)
    c, w1, b1 = r(M, 3) * 4, r(K, 3), r(K)
    ref3 = (c.double() @ w1.double().T + b1.double()).clamp_min(0) @ w.double().T + b.double()
    got3 = ops.rows_linear(c, w, b, mlp3=(w1, b1))
    assert float((got3.double() - ref3).abs().max()) <= 2e-5 * max(1.0, float(ref3.abs().max()))
    # strided rows (a column slice of a wider matrix) and leading batch dimensions
    wide = r(M, K + 64)
    got = ops.rows_linear(wide[:, :K].reshape(1, M, K), w, b)
    assert got.shape == (1, M, N)
    assert float((got[0].double() - (wide[:, :K].double() @ w.double().T + b.double())).abs().max()) <= tol


def test_query_pos_hidden_and_add(dev):
    from pointnet_refine_b200 import ops
    g = torch.Generator(device=dev).manual_seed(5)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    for rows in (1, 32, 1000):
        cur, w1, b1 = r(rows, 3) * 5, r(256, 3), r(256)
        ref = (cur.double() @ w1.double().T + b1.double()).clamp_min(0)
        assert float((ops.query_pos_hidden(w1, b1, cur).double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
        ctx4 = torch.cat([cur, r(rows, 1)], dim=1)                       # context rows [x, y, z, intensity]
        got4 = ops.query_pos_hidden(w1, b1, ctx4, round_tf32=True)
        assert float((got4.double() - ref).abs().max()) <= 2.0 ** -11 * max(1.0, float(ref.abs().max()))
        assert int((got4.view(torch.int32) & 0x1FFF).abs().max()) == 0   # TF32-exact values
        a, b = r(rows, 256), r(rows, 256)
        assert torch.equal(ops.add(a, b), a + b)
        rounded = ops.add(a, None, round_tf32=True)
        assert torch.equal(rounded, _rna_tf32(a))


@pytest.mark.parametrize("B,N,splits", [(1, 1024, None), (2, 700, 3), (1, 65, 1)])
def test_ctx_attention_split_merge(dev, B, N, splits):
    """Splits merged by the native kernel == one unsplit pass (fp32 and bf16 outputs)."""
    from pointnet_refine_b200 import ops
    g = torch.Generator(device=dev).manual_seed(N)
    q = (torch.randn(B, 256, 256, device=dev, generator=g) * 0.3).bfloat16()
    kp = torch.randn(B, N, 256, device=dev, generator=g).bfloat16()
    mem = torch.randn(B, N, 256, device=dev, generator=g).bfloat16()
    one = ops.ctx_attention(q, kp, mem, splits=1)
    s = torch.softmax((q.double() @ kp.double().transpose(1, 2)) * 0.6931471805599453, dim=-1) @ mem.double()
    assert float((one.double() - s).abs().max()) <= 2e-2
    got = ops.ctx_attention(q, kp, mem, splits=splits)
    assert got.dtype == torch.float32 and float((got - one).abs().max()) <= 2e-2
    assert float((got.double() - s).abs().max()) <= 2e-2
    got16 = ops.ctx_attention(q, kp, mem, splits=splits, out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16 and float((got16.float() - got).abs().max()) <= 2.0 ** -7 * max(1.0, float(got.abs().max()))


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
@pytest.mark.parametrize("B,N", [(1, 1024), (2, 300), (7, 513), (9, 300)])
def test_small_batch_forward_vs_live_oracle(dev, B, N, prec):
    """B < 8 (fewer than 256 polyline rows): rows_linear query side, attention split over clusters and merged natively
    (bf16 tier) / fp32 cross attention on hoisted TF32 K / V (tf32 tier); B = 9 takes the tensor-core query side."""
    import pointnet_refine_b200 as prb
    sd = synth.make_state_dict(13)
    ctx, line = synth.make_inputs(B, N, seed=77)
    ref = orc.line_refine_forward(sd, ctx, line)
    m = prb.LineRefineNet().to(dev).eval()
    m.load_state_dict(synth.to_torch(sd), strict=True)
    m.precision = prec
    with torch.no_grad():
        out = m(torch.from_numpy(ctx).to(dev), torch.from_numpy(line).to(dev))
    err = float(np.abs(out.cpu().numpy() - ref).max())
    assert out.shape == (6, B, 32, 3)
    assert err <= (1e-2 * max(1.0, float(np.abs(ref).max())) if prec == "bf16" else 1e-3), err


@pytest.mark.parametrize("B,N,splits", [(1, 200, 0), (3, 1000, 0), (2, 129, 0)])
def test_cross_attention32_matches_torch(dev, B, N, splits):
    """fp32 cross attention of the tf32 tier on a column block of a wider (B, N, 6*256) projection buffer."""
    import torch.nn.functional as F
    from pointnet_refine_b200 import ops
    g = torch.Generator(device=dev).manual_seed(N)
    q = torch.randn(B, 32, 256, device=dev, generator=g)
    kall = torch.randn(B, N, 6 * 256, device=dev, generator=g)
    vall = torch.randn(B, N, 6 * 256, device=dev, generator=g)
    for layer in (0, 5):
        k, v = kall[:, :, layer * 256:(layer + 1) * 256], vall[:, :, layer * 256:(layer + 1) * 256]
        got = ops.cross_attention32(q, k, v)
        heads = lambda t: t.reshape(B, -1, 8, 32).transpose(1, 2).double()
        ref = F.scaled_dot_product_attention(heads(q), heads(k), heads(v)).transpose(1, 2).reshape(B, 32, 256)
        assert float((got.double() - ref).abs().max()) <= 2e-5


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
@pytest.mark.parametrize("B,N", [(1, 1024), (16, 512)])
def test_forward_launches_only_library_kernels(dev, B, N, prec):
    """Eval forward of the bf16 and tf32 tiers, B = 1 whole-scene call and a batch on the tensor-core query side: every
    kernel on the device timeline belongs to this library (lrn:: / scene::); memcpy / memset nodes are allowed."""
    import pointnet_refine_b200 as prb
    from torch.profiler import ProfilerActivity, profile
    sd = synth.make_state_dict(0)
    m = prb.LineRefineNet().to(dev).eval()
    m.load_state_dict(synth.to_torch(sd), strict=True)
    m.precision = prec
    ctx, line = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(B, N, seed=3))
    with torch.no_grad():
        m(ctx, line)                      # weight preparation (host-side folds + uploads) happens on the first call
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m(ctx, line)
            torch.cuda.synchronize()
    names = [e.name for e in prof.events() if "cuda" in str(e.device_type).lower()]
    if not names:
        pytest.skip("the profiler recorded no device activity on this box (CUPTI unavailable)")
    foreign = sorted({n for n in names if not ("lrn::" in n or "scene::" in n or n.lower().startswith(("memcpy", "memset")))})
    assert not foreign, foreign
    assert sum("lrn::" in n for n in names) >= 50
