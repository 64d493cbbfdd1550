"""CPU: pin the numpy oracle against fixtures generated from the unmodified
reference (oracle/make_golden.py).  fp32 vs fp32, different op order -> 1e-5."""
import numpy as np
import pytest

from oracle import lrn_oracle as orc
from tests.golden_util import EVAL_CASES, load_case

TOL = 2e-5


@pytest.mark.parametrize("name", EVAL_CASES)
def test_encoder_matches_reference(name):
    g, sd, ctx, line, (sc, sn) = load_case(name)
    gf, fused, arg = orc.encoder_forward(sd, ctx)
    scale = max(1.0, float(np.abs(g["global_feat"]).max()))
    assert np.abs(gf - g["global_feat"]).max() <= TOL * scale
    fused_ref_layout = fused.transpose(0, 2, 1)            # (B,1024,N) like the reference
    assert np.abs(fused_ref_layout[:, ::sc, ::sn] - g["fused_sub"]).max() <= TOL * scale
    np.testing.assert_allclose(fused.astype(np.float64).sum(1), g["fused_csum"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(fused.astype(np.float64).sum(2), g["fused_psum"], rtol=1e-4, atol=1e-4)
    # argmax: bit-exact wherever the reference's top-2 gap exceeds the fp32 noise
    safe = g["gap"] > 10 * TOL * scale
    assert safe.mean() > 0.5 or ctx.shape[1] == 1
    assert np.array_equal(arg[safe], g["argmax"][safe])


@pytest.mark.parametrize("name", EVAL_CASES)
def test_memory_and_full_forward_match_reference(name):
    g, sd, ctx, line, (sc, sn) = load_case(name)
    _, fused, _ = orc.encoder_forward(sd, ctx)
    mem = orc.context_memory(sd, fused)
    scale = max(1.0, float(np.abs(g["memory_sub"]).max()))
    assert np.abs(mem[:, ::sn, ::sc] - g["memory_sub"]).max() <= TOL * scale
    out = orc.line_refine_forward(sd, ctx, line)
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 2e-4 * max(1.0, float(np.abs(g["out"]).max()))


def test_train_mode_encoder_matches_reference():
    g, sd, ctx, _, (sc, sn) = load_case("train_b2_n512")
    gf, fused, stats = orc.encoder_forward_train(sd, ctx)
    scale = max(1.0, float(np.abs(g["global_feat"]).max()))
    assert np.abs(gf - g["global_feat"]).max() <= 1e-4 * scale
    assert np.abs(fused.transpose(0, 2, 1)[:, ::sc, ::sn] - g["fused_sub"]).max() <= 1e-4 * scale
    for k, v in stats.items():
        ref = g[k.replace(".", "__")]
        np.testing.assert_allclose(np.asarray(v, np.float64), ref, rtol=2e-5, atol=2e-6)


def test_known_answer_head_and_rounding_helpers():
    sd = __import__("oracle.synth", fromlist=["x"]).make_state_dict(0)
    t = np.zeros((1, 1, 256), np.float32)
    d = orc.head_forward(sd, 0, t)
    b1 = np.maximum(sd["reg_branches.0.0.bias"], 0)
    np.testing.assert_allclose(d[0, 0], sd["reg_branches.0.2.weight"] @ b1 + sd["reg_branches.0.2.bias"], rtol=1e-6)
    x = np.array([1.0, 1.00390625, 3.14159265, -2.7182818], np.float32)
    assert np.array_equal(orc.round_bf16(x).view(np.uint32) & 0xFFFF, np.zeros(4, np.uint32))
    assert np.abs(orc.round_bf16(x) - x).max() <= np.abs(x).max() * 2.0 ** -8
    assert np.abs(orc.round_tf32(x) - x).max() <= np.abs(x).max() * 2.0 ** -10


@pytest.mark.parametrize("name", ["b2_n1024", "b3_n1000_ragged", "b2_n2048_realistic"])
def test_torch_port_matches_reference(name):
    """The torch-CPU port used as bench.py's cpu_baseline issues the reference's ops: same numbers."""
    import torch
    from oracle import synth, torch_port
    g, sd, ctx, line, (sc, sn) = load_case(name)
    tsd = synth.to_torch(sd)
    with torch.no_grad():
        gf, fused = torch_port.encoder_forward(tsd, torch.from_numpy(ctx).transpose(2, 1))
        out = torch_port.line_refine_forward(tsd, torch.from_numpy(ctx), torch.from_numpy(line))
    assert np.abs(gf.numpy() - g["global_feat"]).max() <= 1e-6 * max(1.0, float(np.abs(g["global_feat"]).max()))
    assert np.abs(fused.numpy()[:, ::sc, ::sn] - g["fused_sub"]).max() <= 1e-6
    assert np.abs(out.numpy() - g["out"]).max() <= 1e-5
