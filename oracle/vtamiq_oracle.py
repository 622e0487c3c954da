"""TEST INFRASTRUCTURE ONLY — fp32 CPU restatement of ``VTAMIQ.forward`` as a function of the reference
``state_dict`` (same keys).  Written with plain torch CPU ops (the reference's own arithmetic lives in ATen,
SURVEY.md §8c), no nn.Module reuse.  Pinned by tests/golden/forward_*.npz (scores produced by the unmodified
reference, see tests/golden/make_golden.py).

Follows, step by step:
  embeddings ........ modules/VisionTransformer/transformer.py:396-400, :417-426, :507-562
  encoder block ..... transformer.py:153-172, :212-215, :275-285 ; loop + final norm :363-378
  token slice ....... transformer.py:628-641
  VTAMIQ head ....... modules/vtamiq/vtamiq.py:94-119 ; DiffNet modules/RCAN/channel_attention.py:13-86
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

HEAD_DIM = 64
LN_EPS = 1e-6


def _cfg_from_state(sd):
    hidden = sd["transformer.encoder.encoder_norm.weight"].numel()
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.encoder.layers."))
    pos_key = "transformer.embeddings.positional_embeddings.positional_embeddings"
    grid = int(round(math.sqrt(sd[pos_key].shape[1] - 1))) if pos_key in sd else 0
    sc_key = "transformer.embeddings.scale_embeddings.scale_embeddings"
    num_scales = sd[sc_key].shape[1] - 1 if sc_key in sd else 0
    rgs = sorted({int(k.split(".")[1]) for k in sd if k.startswith("quality_decoder.") and ".body." in k})
    n_rcabs = 0
    if rgs:
        n_rcabs = 1 + max(int(k.split(".")[3]) for k in sd
                          if k.startswith("quality_decoder.0.body.") and k.split(".")[4] == "body")
    return dict(hidden=hidden, heads=hidden // HEAD_DIM, layers=n_layers, grid=grid, num_scales=num_scales,
                num_rgs=len(rgs), num_rcabs=n_rcabs)


def embed(sd, cfg, x, pos, scales):
    """x (B,N,3,P,P) or (B,N,H) -> (B, T+N, H)."""
    e = "transformer.embeddings."
    if x.dim() == 5:
        B, N = x.shape[:2]
        x = F.conv2d(x.reshape(B * N, *x.shape[2:]), sd[e + "patch_embeddings.weight"],
                     sd[e + "patch_embeddings.bias"], stride=x.shape[-1])
    else:
        B, N = x.shape[:2]
    x = x.reshape(B, N, -1)
    if cfg["grid"]:
        table = sd[e + "positional_embeddings.positional_embeddings"]
        g = cfg["grid"]
        p = torch.floor(pos.reshape(B * N, 2) * g)
        idx = ((p[:, 0] * g + p[:, 1]) + 1).to(torch.long)
        x = x + table[:, idx].reshape(B, N, -1)
    if cfg["num_scales"]:
        if scales is None:
            raise ValueError("Model uses scale embedding but scales is passed as None.")
        sidx = (torch.clamp(scales.reshape(B * N), 0, cfg["num_scales"] - 1) + 1).to(torch.long)
        x = x + sd[e + "scale_embeddings.scale_embeddings"][:, sidx].reshape(B, N, -1)
    toks = []
    if e + "cls_token" in sd:
        cls = sd[e + "cls_token"].expand(B, 1, -1)
        if cfg["grid"]:
            cls = cls + sd[e + "positional_embeddings.positional_embeddings"][:, 0]
        toks.append(cls)
    if e + "extra_tokens" in sd:
        toks.append(sd[e + "extra_tokens"].expand(B, -1, -1))
    if toks:
        x = torch.cat(toks + [x], dim=1)
    return x


def _adapter(sd, p, h):
    """Houlsby adapter h + W2 gelu(W1 h + b1) + b2 (transformer.py:177-194); the VTAMIQ path always uses adapter 0 of a
    layer when the model has adapters (backbone.py:54-59).  No-op when the layer has none."""
    if p + "0.weight" not in sd:
        return h
    return h + F.linear(F.gelu(F.linear(h, sd[p + "0.weight"], sd[p + "0.bias"])), sd[p + "2.weight"], sd[p + "2.bias"])


def encoder_layer(sd, cfg, x, i):
    p = f"transformer.encoder.layers.{i}."
    B, S, H = x.shape
    nh = cfg["heads"]
    y = F.layer_norm(x, (H,), sd[p + "attention_norm.weight"], sd[p + "attention_norm.bias"], LN_EPS)
    split = lambda t: t.view(B, S, nh, HEAD_DIM).permute(0, 2, 1, 3)
    q = split(F.linear(y, sd[p + "attn.query.weight"], sd[p + "attn.query.bias"]))
    k = split(F.linear(y, sd[p + "attn.key.weight"], sd[p + "attn.key.bias"]))
    v = split(F.linear(y, sd[p + "attn.value.weight"], sd[p + "attn.value.bias"]))
    prob = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(HEAD_DIM), dim=-1)
    a = torch.matmul(prob, v).permute(0, 2, 1, 3).contiguous().view(B, S, H)
    a = F.linear(a, sd[p + "attn.out.weight"], sd[p + "attn.out.bias"])
    a = _adapter(sd, p + "adapter1.adapter.", a)
    if p + "ls1.gamma" in sd:
        a = a * sd[p + "ls1.gamma"]
    x = x + a
    y = F.layer_norm(x, (H,), sd[p + "ffn_norm.weight"], sd[p + "ffn_norm.bias"], LN_EPS)
    y = F.gelu(F.linear(y, sd[p + "ffn.fc1.weight"], sd[p + "ffn.fc1.bias"]))
    y = F.linear(y, sd[p + "ffn.fc2.weight"], sd[p + "ffn.fc2.bias"])
    y = _adapter(sd, p + "adapter2.adapter.", y)
    if p + "ls2.gamma" in sd:
        y = y * sd[p + "ls2.gamma"]
    return x + y


def vit_tokens(sd, cfg, patches, pos, scales, return_layers=False):
    """ViT forward -> all tokens after encoder_norm (B, S, H); optionally the residual stream after each block."""
    x = embed(sd, cfg, patches, pos, scales)
    states = [x]
    for i in range(cfg["layers"]):
        x = encoder_layer(sd, cfg, x, i)
        states.append(x)
    H = x.shape[-1]
    out = F.layer_norm(x, (H,), sd["transformer.encoder.encoder_norm.weight"],
                       sd["transformer.encoder.encoder_norm.bias"], LN_EPS)
    return (out, states) if return_layers else out


def diffnet_head(sd, cfg, d, drop_scale=None):
    """d (B,H) = gamma*(cls_ref - cls_dist) -> q (B,).  1x1 Conv1d on a length-1 signal == F.linear.
    drop_scale (num_rgs, B) or None: training-mode DropPath of each ResidualGroup branch as the per-pair factor
    mask/keep_prob (channel_attention.py:26-29 ``x + self.drop(self.body(x))``; timm DropPath).  Differentiable:
    tests run it under torch.autograd to check the CUDA backward."""
    lin = lambda x, w, b: F.linear(x, sd[w].squeeze(-1), sd[b])
    x = d
    for g in range(cfg["num_rgs"]):
        skip = x
        for r in range(cfg["num_rcabs"]):
            p = f"quality_decoder.{g}.body.{r}.body."
            y = lin(F.prelu(x, sd[p + "1.weight"]), p + "2.weight", p + "2.bias")
            w = torch.relu(lin(y, p + "4.conv_du.1.weight", p + "4.conv_du.1.bias"))
            w = torch.sigmoid(lin(w, p + "4.conv_du.4.weight", p + "4.conv_du.4.bias"))
            x = x + y * w
        p = f"quality_decoder.{g}.body.{cfg['num_rcabs']}."
        branch = lin(x, p + "weight", p + "bias")
        if drop_scale is not None:
            branch = branch * drop_scale[g][:, None]
        x = skip + branch
    if cfg["num_rgs"]:
        p = f"quality_decoder.{cfg['num_rgs']}."
        x = lin(x, p + "weight", p + "bias")
    x = F.linear(x, sd["q_predictor.1.weight"], sd["q_predictor.1.bias"])
    x = F.prelu(x, sd["q_predictor.2.weight"])
    x = F.linear(x, sd["q_predictor.4.weight"], sd["q_predictor.4.bias"])
    return x.flatten()


@torch.no_grad()
def vtamiq_forward(sd, patches, pos, scales, token_num=0, return_intermediates=False):
    """sd: reference state_dict (CPU fp32). patches/pos/scales: (ref, dist) tuples as the model receives them."""
    cfg = _cfg_from_state(sd)
    feats = []
    inter = {}
    for i in range(2):
        sc = scales[i] if scales is not None else None
        if return_intermediates:
            f, states = vit_tokens(sd, cfg, patches[i], pos[i], sc, return_layers=True)
            inter[("ref", "dist")[i]] = states
        else:
            f = vit_tokens(sd, cfg, patches[i], pos[i], sc)
        feats.append(f[:, token_num])
    d = feats[0] - feats[1]
    if "diff_scale.gamma" in sd:
        d = d * sd["diff_scale.gamma"]
    q = diffnet_head(sd, cfg, d)
    if return_intermediates:
        inter["diff"] = d
        return q, inter
    return q
