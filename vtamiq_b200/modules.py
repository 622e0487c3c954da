"""Parameter containers that mirror the reference module tree for the VTAMIQ hot path.

These classes exist for ONE reason: the drop-in boundary requires the new
``VTAMIQ`` to expose the same ``state_dict`` keys/shapes, the same attribute paths that
``train.py`` / ``backbone.py`` touch, and — because the parity fixtures are seeded — the same
random-init stream as the reference when built after ``torch.manual_seed(s)``.

None of the ``nn.Linear`` / ``nn.Conv*`` objects below is ever *called*: the forward pass lives
in :mod:`vtamiq_b200.engine` and runs hand-written sm_100a kernels through the C-ABI library.
There is no torch/CPU forward in this package.

Reference layout being mirrored (file:line under the reference repo):
  * ViT configs ............... modules/VisionTransformer/transformer.py:68-111
  * attention / MLP params .... transformer.py:125-146, :197-210
  * encoder layer / encoder ... transformer.py:246-273, :328-361
  * embeddings ................ transformer.py:385-394, :403-415, :458-505
  * ViT wrapper + init ........ transformer.py:565-626, :671-678
  * npz loader ................ transformer.py:287-325, :428-455, :643-668
  * DiffNet (RCAN blocks) ..... modules/RCAN/channel_attention.py:13-86, modules/vtamiq/vtamiq.py:12-23
  * quality head .............. modules/vtamiq/vtamiq.py:71-77
"""
from __future__ import annotations

import math
import warnings
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

VIT_VARIANT_B8 = "ViT-B8"
VIT_VARIANT_B16 = "ViT-B16"
VIT_VARIANT_L16 = "ViT-L16"

_TOKEN_INIT_STD = 0.02

_WEIGHTS_DIR = "./modules/VisionTransformer/weights/"


def get_vit_config(variant: str) -> OrderedDict:
    """Architecture table; keys match transformer.py:68-111."""
    table = {
        VIT_VARIANT_B16: ("imagenet21k+imagenet2012_ViT-B_16.npz", 16, 768, 3072, 12, 12),
        VIT_VARIANT_B8: ("imagenet21k+imagenet2012_ViT-B_8.npz", 8, 768, 3072, 12, 12),
        VIT_VARIANT_L16: ("imagenet21k+imagenet2012_ViT-L_16.npz", 16, 1024, 4096, 16, 24),
    }
    if variant not in table:
        raise ValueError("ViT: Unsupported variant [{}], pick from {}.".format(variant, list(table)))
    fname, patch, hidden, mlp, heads, layers = table[variant]
    return OrderedDict(
        vit_weights_path=_WEIGHTS_DIR + fname,
        img_dim=384,
        patch_size=patch,
        hidden_size=hidden,
        mlp_dim=mlp,
        num_heads=heads,
        num_layers=layers,
    )


def _warn_unused(tag, kwargs):
    # reference: utils/misc/miscelaneous.py:8-10 — unknown kwargs only warn
    for k, v in kwargs.items():
        warnings.warn(f"{tag}: Unused kwarg [{k}={v}]")


def _npz_tensor(a: np.ndarray) -> torch.Tensor:
    """HWIO conv kernels become OIHW; everything else is taken as is (transformer.py:118-122)."""
    if a.ndim == 4:
        a = a.transpose(3, 2, 0, 1)
    return torch.from_numpy(np.ascontiguousarray(a))


class LayerScale(nn.Module):
    """Per-channel gain, ``gamma`` (transformer.py:235-243)."""

    def __init__(self, dim, init_values=1.0):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class MultiHeadSelfAttention(nn.Module):
    """query/key/value/out projections (transformer.py:125-146). Parameters only."""

    def __init__(self, hidden, heads):
        super().__init__()
        self.num_attention_heads = heads
        self.attention_head_size = hidden // heads
        self.all_head_size = heads * self.attention_head_size
        self.query = nn.Linear(hidden, self.all_head_size)
        self.key = nn.Linear(hidden, self.all_head_size)
        self.value = nn.Linear(hidden, self.all_head_size)
        self.out = nn.Linear(hidden, hidden)


class MLP(nn.Module):
    """fc1/fc2 (transformer.py:197-210); xavier+tiny-bias init is consumed from the RNG stream
    in the same order as the reference even though ``_init_weights`` overwrites it later."""

    def __init__(self, hidden, mlp_dim):
        super().__init__()
        self.fc1 = nn.Linear(hidden, mlp_dim)
        self.fc2 = nn.Linear(mlp_dim, hidden)
        for fc in (self.fc1, self.fc2):
            nn.init.xavier_uniform_(fc.weight)
            nn.init.normal_(fc.bias, std=1e-6)


class Adapter(nn.Module):
    """Houlsby adapter parameters: channels -> channels/4 -> channels (transformer.py:177-194); the constructor draws
    from the RNG exactly like the reference's (nn.Linear defaults, then xavier + tiny-bias)."""

    def __init__(self, channels, reduction=4):
        super().__init__()
        self.adapter = nn.Sequential(nn.Linear(channels, channels // reduction), nn.GELU(),
                                     nn.Linear(channels // reduction, channels))
        for layer in self.adapter:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)
                nn.init.normal_(layer.bias, std=1e-6)


class EncoderLayer(nn.Module):
    """One pre-LN block's parameters (transformer.py:246-273)."""

    def __init__(self, cfg, use_layer_scale, num_adapters):
        super().__init__()
        hidden = cfg["hidden_size"]
        self.hidden_size = hidden
        self.attention_norm = nn.LayerNorm(hidden, eps=1e-6)
        self.ffn_norm = nn.LayerNorm(hidden, eps=1e-6)
        self.ffn = MLP(hidden, cfg["mlp_dim"])
        self.attn = MultiHeadSelfAttention(hidden, cfg["num_heads"])
        self.use_adapters = num_adapters > 0
        if self.use_adapters:   # same module names / construction order as transformer.py:258-267
            self.adapters = []
            for i in range(num_adapters):
                a1, a2 = Adapter(hidden), Adapter(hidden)
                self.add_module(f"adapter{2 * i + 1}", a1)
                self.add_module(f"adapter{2 * i + 2}", a2)
                self.adapters.append((a1, a2))
        self.ls1 = LayerScale(hidden) if use_layer_scale else nn.Identity()
        self.ls2 = LayerScale(hidden) if use_layer_scale else nn.Identity()

    @torch.no_grad()
    def load_from(self, weights, n_block):
        """JAX checkpoint → torch layout for one block (transformer.py:287-325)."""
        root = f"Transformer/encoderblock_{n_block}/"
        h = self.hidden_size
        att = root + "MultiHeadDotProductAttention_1/"
        for name, lin in (("query", self.attn.query), ("key", self.attn.key),
                          ("value", self.attn.value), ("out", self.attn.out)):
            lin.weight.copy_(_npz_tensor(weights[att + name + "/kernel"]).view(h, h).t())
            lin.bias.copy_(_npz_tensor(weights[att + name + "/bias"]).view(-1))
        for name, lin in (("Dense_0", self.ffn.fc1), ("Dense_1", self.ffn.fc2)):
            lin.weight.copy_(_npz_tensor(weights[root + "MlpBlock_3/" + name + "/kernel"]).t())
            lin.bias.copy_(_npz_tensor(weights[root + "MlpBlock_3/" + name + "/bias"]).t())
        for name, ln in (("LayerNorm_0", self.attention_norm), ("LayerNorm_2", self.ffn_norm)):
            ln.weight.copy_(_npz_tensor(weights[root + name + "/scale"]))
            ln.bias.copy_(_npz_tensor(weights[root + name + "/bias"]))


class Encoder(nn.Module):
    """encoder_norm + ``layers`` ModuleList (transformer.py:328-361). ``layers`` must stay a sized,
    iterable container: train.py:691 takes its len, backbone.py:73-82 iterates it."""

    def __init__(self, cfg, num_keep_layers, num_adapters, use_layer_scale):
        super().__init__()
        n = cfg["num_layers"]
        if num_keep_layers > 0:
            n = max(1, min(num_keep_layers, cfg["num_layers"]))
        self.encoder_norm = nn.LayerNorm(cfg["hidden_size"], eps=1e-6)
        self.layers = nn.ModuleList()
        self.num_layers = n
        for _ in range(n):
            self.layers.append(EncoderLayer(cfg, use_layer_scale, num_adapters))


class ScaleEmbedding(nn.Module):
    """(1, num_scales+1, H) table, row 0 unused (transformer.py:385-394)."""

    def __init__(self, num_scales, hidden):
        super().__init__()
        self.num_scales = num_scales
        self.scale_embeddings = nn.Parameter(torch.zeros(1, num_scales + 1, hidden))
        self.scale_embeddings.data.normal_(mean=0.0, std=_TOKEN_INIT_STD)


class UvPosEmbedding(nn.Module):
    """(1, (img_dim/patch)^2+1, H) table indexed by uv (transformer.py:403-415)."""

    def __init__(self, cfg):
        super().__init__()
        self.width_pos_embeddings = cfg["img_dim"] // cfg["patch_size"]
        n = self.width_pos_embeddings ** 2 + 1
        self.positional_embeddings = nn.Parameter(torch.zeros(1, n, cfg["hidden_size"]))
        self.positional_embeddings.data.normal_(mean=0.0, std=_TOKEN_INIT_STD)

    @torch.no_grad()
    def load_from(self, weights):
        """transformer.py:428-455: copy as-is when shapes agree, else bilinear grid resize."""
        from scipy import ndimage
        posemb = _npz_tensor(weights["Transformer/posembed_input/pos_embedding"])
        if posemb.size() != self.positional_embeddings.size():
            ntok_new = self.positional_embeddings.size(1) - 1
            tok, grid = posemb[:, :1], posemb[0, 1:]
            gs_old, gs_new = int(np.sqrt(len(grid))), int(np.sqrt(ntok_new))
            grid = grid.reshape(gs_old, gs_old, -1)
            z = gs_new / gs_old
            grid = ndimage.zoom(grid, (z, z, 1), order=1).reshape(1, gs_new * gs_new, -1)
            posemb = _npz_tensor(np.concatenate([tok, grid], axis=1))
        self.positional_embeddings.copy_(posemb)


class Embeddings(nn.Module):
    """patch projection, tokens, pos/scale tables (transformer.py:458-505). Parameters only."""

    def __init__(self, cfg, use_cls_token, use_patch_embedding, use_pos_embedding,
                 num_extra_tokens, num_scales):
        super().__init__()
        hidden, patch = cfg["hidden_size"], cfg["patch_size"]
        self.use_patch_embedding = use_patch_embedding
        if use_patch_embedding:
            self.patch_embeddings = nn.Conv2d(3, hidden, kernel_size=patch, stride=patch)
        self.use_cls_token = use_cls_token
        if use_cls_token:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden), requires_grad=True)
        self.cls_token.data.normal_(mean=0.0, std=_TOKEN_INIT_STD)  # raises w/o cls, like the reference
        self.num_extra_tokens = num_extra_tokens
        self.use_extra_tokens = num_extra_tokens > 0
        if self.use_extra_tokens:
            self.extra_tokens = nn.Parameter(torch.zeros(1, num_extra_tokens, hidden), requires_grad=True)
            self.extra_tokens.data.normal_(mean=0.0, std=_TOKEN_INIT_STD)
        self.use_tokens = self.use_cls_token or self.use_extra_tokens
        self.num_tokens = int(use_cls_token) + num_extra_tokens
        self.use_pos_embedding = use_pos_embedding
        if use_pos_embedding:
            self.positional_embeddings = UvPosEmbedding(cfg)
        self.use_scale_embedding = num_scales > 1
        if self.use_scale_embedding:
            self.scale_embeddings = ScaleEmbedding(num_scales, hidden)


class VisionTransformer(nn.Module):
    """Embeddings + Encoder parameter tree and weight IO (transformer.py:565-678)."""

    def __init__(self, config, use_patch_embedding=True, use_pos_embedding=True, use_cls_token=True,
                 use_classifier=False, use_layer_scale=False, num_keep_layers=-1, num_extra_tokens=0,
                 num_classes=1000, num_adapters=0, num_scales=0, path_drop_prob=0., pretrained=True,
                 return_layers=False, return_attention=False):
        super().__init__()
        if use_classifier:
            raise NotImplementedError("vtamiq_b200: the ImageNet classifier head is not on the VTAMIQ path")
        self.config = config
        self.hidden_size = config["hidden_size"]
        self.use_cls_token = use_cls_token
        self.num_extra_tokens = num_extra_tokens
        self.use_extra_tokens = num_extra_tokens > 0
        self.use_classifier = False
        self.use_layer_scale = use_layer_scale
        self.use_adapters = num_adapters > 0
        self.embeddings = Embeddings(config, use_cls_token, use_patch_embedding, use_pos_embedding,
                                     num_extra_tokens, num_scales)
        self.use_tokens = self.use_cls_token or self.use_extra_tokens
        self.num_tokens = self.embeddings.num_tokens
        self.encoder = Encoder(config, num_keep_layers, num_adapters, use_layer_scale)
        if pretrained:
            path = config["vit_weights_path"]
            print("ViT: Loading pretrained transformer from path:", path)
            self.load_from(np.load(path), use_patch_embedding, use_pos_embedding)
        else:
            self.apply(self._init_weights)

    @torch.no_grad()
    def load_from(self, weights, use_patch_embedding=True, use_pos_embedding=True):
        """ViT-*.npz → parameters; same key/transposition rules as transformer.py:643-668."""
        emb = self.embeddings
        if use_patch_embedding:
            emb.patch_embeddings.weight.copy_(_npz_tensor(weights["embedding/kernel"]))
            emb.patch_embeddings.bias.copy_(_npz_tensor(weights["embedding/bias"]))
        if self.use_tokens and self.use_cls_token:
            emb.cls_token.copy_(_npz_tensor(weights["cls"]))
        if use_pos_embedding:
            emb.positional_embeddings.load_from(weights)
        self.encoder.encoder_norm.weight.copy_(_npz_tensor(weights["Transformer/encoder_norm/scale"]))
        self.encoder.encoder_norm.bias.copy_(_npz_tensor(weights["Transformer/encoder_norm/bias"]))
        for idx, layer in enumerate(self.encoder.layers):
            layer.load_from(weights, n_block=str(idx))

    @staticmethod
    def _init_weights(m):
        # DeiT-style init for pretrained=False (transformer.py:671-678)
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=_TOKEN_INIT_STD)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


# --------------------------------------------------------------------------------------------
# DiffNet containers.  Index positions inside each nn.Sequential are part of the state_dict
# contract (…body.{r}.body.1.weight, …body.4.conv_du.{1,4}.*, quality_decoder.{g}.body.4.*).
# --------------------------------------------------------------------------------------------

class CALayer(nn.Module):
    """Channel attention squeeze/excite convs (channel_attention.py:53-86); on a length-1 signal
    the pooling at index 0 is the identity, kept only so the conv indices stay 1 and 4."""

    def __init__(self, dim, reduction):
        super().__init__()
        hidden = dim // reduction
        self.conv_du = nn.Sequential(
            nn.AdaptiveAvgPool1d(1),
            nn.Conv1d(dim, hidden, kernel_size=1),
            nn.Sequential(),
            nn.ReLU(inplace=True),
            nn.Conv1d(hidden, dim, kernel_size=1),
            nn.Sequential(),
        )


class RCAB(nn.Module):
    """[-, PReLU, Conv1d, -, CALayer] (channel_attention.py:34-50 with use_bn=False)."""

    def __init__(self, dim, reduction):
        super().__init__()
        self.body = nn.Sequential(
            nn.Sequential(),
            nn.PReLU(),
            nn.Conv1d(dim, dim, kernel_size=1),
            nn.Sequential(),
            CALayer(dim, reduction),
        )


class ResidualGroup(nn.Module):
    """``num_rcabs`` RCABs then a 1x1 conv, wrapped by a skip (channel_attention.py:13-29)."""

    def __init__(self, dim, num_rcabs, reduction):
        super().__init__()
        self.body = nn.Sequential(*[RCAB(dim, reduction) for _ in range(num_rcabs)],
                                  nn.Conv1d(dim, dim, kernel_size=1))


def make_quality_decoder(dim, num_rgs, num_rcabs, ca_reduction):
    """vtamiq.py:12-23."""
    return nn.Sequential(*[ResidualGroup(dim, num_rcabs, ca_reduction) for _ in range(num_rgs)],
                         nn.Conv1d(dim, dim, kernel_size=1))


def set_grad(layer, requires_grad):
    for p in layer.parameters():
        p.requires_grad = requires_grad
