"""Deterministic synthetic weights and inputs for parity tests and benchmarks.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything is drawn from
``numpy.random.Generator(PCG64(seed))`` whose stream is stable across numpy
releases, so the GPU box regenerates bit-identical weights/inputs from a seed
and only the (small) reference *outputs* have to be committed as fixtures.

The parameter tree is the reference's 205-key ``state_dict``
(/root/reference/src/model.py:7-37,138-179; listing in SURVEY.md appendix B).
Weight scales follow PyTorch's default initialisers (uniform +-1/sqrt(fan_in));
BatchNorm affine/running statistics are randomised so BN folding is exercised
(SURVEY.md section 8d "Synthetic inputs").
"""
from __future__ import annotations

import numpy as np

ENC_CHANNELS = (4, 64, 128, 256, 512, 1024)   # src/model.py:10-14
FUSION_IN = 64 + 128 + 256 + 512 + 1024       # src/model.py:24
D_MODEL = 256                                 # src/model.py:142
N_LAYERS = 6                                  # src/model.py:143
FFN = 1024                                    # src/model.py:166


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _conv(rng, sd, name, cout, cin, kernel_dim=True):
    bound = 1.0 / np.sqrt(cin)
    shape = (cout, cin, 1) if kernel_dim else (cout, cin)
    sd[name + ".weight"] = _uniform(rng, shape, bound)
    sd[name + ".bias"] = _uniform(rng, (cout,), bound)


def _bn(rng, sd, name, c, randomise):
    if randomise:
        sd[name + ".weight"] = rng.uniform(0.5, 1.5, size=(c,)).astype(np.float32)
        sd[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[name + ".running_mean"] = (0.2 * rng.standard_normal(c)).astype(np.float32)
        sd[name + ".running_var"] = rng.uniform(0.5, 1.5, size=(c,)).astype(np.float32)
    else:
        sd[name + ".weight"] = np.ones(c, np.float32)
        sd[name + ".bias"] = np.zeros(c, np.float32)
        sd[name + ".running_mean"] = np.zeros(c, np.float32)
        sd[name + ".running_var"] = np.ones(c, np.float32)
    sd[name + ".num_batches_tracked"] = np.array(0, dtype=np.int64)


def make_state_dict(seed: int = 0, randomise_bn: bool = True) -> dict:
    """Full LineRefineNet state_dict (205 keys) as numpy arrays, in the
    reference's registration order (src/model.py:138-179)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: dict = {}
    enc = "context_encoder."
    for k in range(1, 6):
        _conv(rng, sd, f"{enc}conv{k}", ENC_CHANNELS[k], ENC_CHANNELS[k - 1])
    for k in range(1, 6):
        _bn(rng, sd, f"{enc}bn{k}", ENC_CHANNELS[k], randomise_bn)
    _conv(rng, sd, f"{enc}fusion.0", 1024, FUSION_IN)
    _bn(rng, sd, f"{enc}fusion.1", 1024, randomise_bn)
    _conv(rng, sd, f"{enc}intensity_gate.0", 64, 1)
    _conv(rng, sd, f"{enc}intensity_gate.2", 1024, 64)
    _conv(rng, sd, "context_proj", D_MODEL, 1024, kernel_dim=False)
    for idx, (cout, cin) in zip((0, 3, 6), ((64, 3), (128, 64), (D_MODEL, 128))):
        _conv(rng, sd, f"point_mlp.{idx}", cout, cin)
        _bn(rng, sd, f"point_mlp.{idx + 1}", cout, randomise_bn)
    _conv(rng, sd, "pos_emb.mlp.0", D_MODEL, 3, kernel_dim=False)
    _conv(rng, sd, "pos_emb.mlp.2", D_MODEL, D_MODEL, kernel_dim=False)
    for l in range(N_LAYERS):
        p = f"decoder_layers.{l}."
        for attn in ("self_attn", "cross_attn"):
            b = np.sqrt(6.0 / (3 * D_MODEL + D_MODEL))  # xavier_uniform on (768,256)
            sd[f"{p}{attn}.in_proj_weight"] = _uniform(rng, (3 * D_MODEL, D_MODEL), b)
            sd[f"{p}{attn}.in_proj_bias"] = (0.02 * rng.standard_normal(3 * D_MODEL)).astype(np.float32)
            _conv(rng, sd, f"{p}{attn}.out_proj", D_MODEL, D_MODEL, kernel_dim=False)
        _conv(rng, sd, f"{p}linear1", FFN, D_MODEL, kernel_dim=False)
        _conv(rng, sd, f"{p}linear2", D_MODEL, FFN, kernel_dim=False)
        for n in (1, 2, 3):
            sd[f"{p}norm{n}.weight"] = rng.uniform(0.8, 1.2, size=(D_MODEL,)).astype(np.float32)
            sd[f"{p}norm{n}.bias"] = (0.05 * rng.standard_normal(D_MODEL)).astype(np.float32)
    for l in range(N_LAYERS):
        _conv(rng, sd, f"reg_branches.{l}.0", 128, D_MODEL, kernel_dim=False)
        _conv(rng, sd, f"reg_branches.{l}.2", 3, 128, kernel_dim=False)
    assert len(sd) == 205, len(sd)
    return sd


def make_inputs(B: int, N: int, M: int = 32, seed: int = 1234, dist: str = "parity"):
    """(context (B,N,4), noisy_line (B,M,3)) float32.

    ``parity``   : N(0,1) everywhere, like the reference's own smoke test
                   (src/model.py:238-239).
    ``realistic``: metres-scale lane crops with integer intensity counts
                   (src/dataset.py:214-237, tools/augment_train_data.py:23-48).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if dist == "parity":
        ctx = rng.standard_normal((B, N, 4), dtype=np.float32)
        line = rng.standard_normal((B, M, 3), dtype=np.float32)
        return ctx, line
    if dist == "realistic":
        ctx = np.empty((B, N, 4), np.float32)
        ctx[..., 0] = rng.uniform(-25.0, 25.0, size=(B, N))
        ctx[..., 1] = 0.25 * rng.standard_normal((B, N))
        ctx[..., 2] = 0.1 * rng.standard_normal((B, N))
        ctx[..., 3] = np.clip(np.round(rng.gamma(2.0, 8.0, size=(B, N))), 0, 255)
        t = np.linspace(-25.0, 25.0, M, dtype=np.float32)
        line = np.zeros((B, M, 3), np.float32)
        line[..., 0] = t[None, :]
        line += (0.05 * rng.standard_normal((B, M, 3))).astype(np.float32)
        line[..., 1] += rng.uniform(-0.4, 0.4, size=(B, 1)).astype(np.float32)
        return ctx, line
    raise ValueError(dist)


def to_torch(sd: dict):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).clone() for k, v in sd.items()}
