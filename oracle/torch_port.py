"""torch-CPU port of the reference forward, op for op (CPU baseline for bench.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is pure PyTorch and cannot
travel to the GPU box (/root/reference is absent there), so the timed CPU baseline is this
functional restatement: it issues the same ATen ops, on the same shapes and in the same order
as /root/reference/src/model.py:39-62 and :181-234 (Conv1d k=1 -> batch_norm -> relu, cat,
sigmoid gate, max/mean; context_proj; pos_emb; point_mlp; 6 x DETR decoder layer + heads), driven
by a plain state_dict.  kind = "port" in bench.py's cpu_baseline.  Checked against the golden
fixtures in tests/test_oracle.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

ENC = "context_encoder."


def _conv_bn(sd, x, conv, bn, relu=True):
    y = F.conv1d(x, sd[conv + ".weight"], sd[conv + ".bias"])
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     False, 0.1, 1e-5)
    return F.relu(y) if relu else y


def encoder_forward(sd, x):
    """x (B,4,N) -> (global_feat (B,2048), fused (B,1024,N)); src/model.py:39-62, eval mode."""
    intensity = x[:, 3:4, :]
    feats, h = [], x
    for k in range(1, 6):
        h = _conv_bn(sd, h, f"{ENC}conv{k}", f"{ENC}bn{k}")
        feats.append(h)
    fused = _conv_bn(sd, torch.cat(feats, dim=1), f"{ENC}fusion.0", f"{ENC}fusion.1")
    g = F.relu(F.conv1d(intensity, sd[f"{ENC}intensity_gate.0.weight"], sd[f"{ENC}intensity_gate.0.bias"]))
    g = torch.sigmoid(F.conv1d(g, sd[f"{ENC}intensity_gate.2.weight"], sd[f"{ENC}intensity_gate.2.bias"]))
    fused = fused * (0.5 + 0.5 * g)
    gf = torch.cat([torch.max(fused, 2)[0], torch.mean(fused, 2)], dim=1)
    return gf, fused


def _pos_emb(sd, xyz):
    h = F.relu(F.linear(xyz, sd["pos_emb.mlp.0.weight"], sd["pos_emb.mlp.0.bias"]))
    return F.linear(h, sd["pos_emb.mlp.2.weight"], sd["pos_emb.mlp.2.bias"])


def _mha(sd, p, q, k, v):
    out, _ = F.multi_head_attention_forward(
        q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), 256, 8,
        sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"], None, None, False, 0.0,
        sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"], training=False, need_weights=True)
    return out.transpose(0, 1)


def line_refine_forward(sd, context, noisy_line):
    """LineRefineNet.forward, eval mode (src/model.py:181-234) -> (6,B,M,3)."""
    _, fused = encoder_forward(sd, context.transpose(2, 1))
    memory = F.linear(fused.transpose(2, 1), sd["context_proj.weight"], sd["context_proj.bias"])
    pos_mem = _pos_emb(sd, context[:, :, :3])
    t = noisy_line.transpose(2, 1)
    for idx, relu in ((0, True), (3, True), (6, False)):
        t = _conv_bn(sd, t, f"point_mlp.{idx}", f"point_mlp.{idx + 1}", relu)
    tgt = t.transpose(2, 1)
    cur = noisy_line.clone()
    outs = []
    for l in range(6):
        p = f"decoder_layers.{l}."
        ln = lambda x, i: F.layer_norm(x, (256,), sd[f"{p}norm{i}.weight"], sd[f"{p}norm{i}.bias"], 1e-5)
        qpos = _pos_emb(sd, cur)
        q = tgt + qpos
        tgt = ln(tgt + _mha(sd, p + "self_attn", q, q, tgt), 1)
        tgt = ln(tgt + _mha(sd, p + "cross_attn", tgt + qpos, memory + pos_mem, memory), 2)
        ff = F.linear(F.relu(F.linear(tgt, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                      sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        tgt = ln(tgt + ff, 3)
        r = f"reg_branches.{l}."
        d = F.linear(F.relu(F.linear(tgt, sd[r + "0.weight"], sd[r + "0.bias"])), sd[r + "2.weight"], sd[r + "2.bias"])
        cur = cur + d
        outs.append(cur - noisy_line)
    return torch.stack(outs)
