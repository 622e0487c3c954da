"""Host-side orchestration of the sm_100a kernels for one VTAMIQ forward.

PyTorch is used here for device memory, streams and CUDA-graph capture only; every arithmetic step
of the path is a call into libvtamiq_b200.so (see include/vtamiq_b200.h).  The call sequence
restates reference ``VTAMIQ.forward`` (modules/vtamiq/vtamiq.py:94-119) →
``VisionTransformer.forward`` (modules/VisionTransformer/transformer.py:628-641) →
``Embeddings.forward`` (:526-562) → 12 × ``EncoderLayer.forward`` (:275-285) → ``encoder_norm``
(:376) → DiffNet + head, with ref and dist stacked into ONE pass of 2B sequences.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (EPI_BIAS_F32, EPI_BIAS_GELU_H, EPI_BIAS_H, EPI_BIAS_RESID_F32, VTQ_BF16, VTQ_F16,
                   VtqError, get_context)

_DTYPES = {"fp16": (VTQ_F16, torch.float16), "bf16": (VTQ_BF16, torch.bfloat16)}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device=None):
    """The caller's current stream ON `device` (not on whatever device happens to be current)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class _LayerPack:
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    w_qkv: torch.Tensor
    b_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    g1: torch.Tensor | None
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    w_fc1: torch.Tensor
    b_fc1: torch.Tensor
    w_fc2: torch.Tensor
    b_fc2: torch.Tensor
    g2: torch.Tensor | None
    # LayerNorm folded into the consuming GEMM (vtq_gemm_ln): W' = W * ln_w, b' = b + W ln_b, colsum = sum_k W'
    # — built only when fuse_layernorm is on
    w_qkv_f: torch.Tensor | None = None
    b_qkv_f: torch.Tensor | None = None
    cs_qkv: torch.Tensor | None = None
    w_fc1_f: torch.Tensor | None = None
    b_fc1_f: torch.Tensor | None = None
    cs_fc1: torch.Tensor | None = None
    # Houlsby adapters behind the attention / MLP sub-blocks: (W1 16-bit, b1, W2 16-bit, b2, width) or None
    ad1: tuple | None = None
    ad2: tuple | None = None


def fold_layernorm(w, b, ln_w, ln_b, dtype16):
    """Operands of vtq_gemm_ln's consumer side for ``Linear(LayerNorm(x))``.

    LN(x) W^T + b == rstd * (x W'^T - mean * colsum) + b'   with   W' = W * ln_w (rounded to the 16-bit operand type),
    b' = b + W ln_b,  colsum = sum_k W'[n, k] (of the ROUNDED W', the matrix the tensor core actually multiplies).
    Returns (W' 16-bit, b' fp32, colsum fp32).
    """
    w32 = w.detach().float()
    wf = (w32 * ln_w.detach().float()[None, :]).to(dtype16).contiguous()
    bf = (b.detach().float() + w32 @ ln_b.detach().float()).contiguous()
    return wf, bf, wf.float().sum(1).contiguous()


class _Workspace:
    """Device buffers for one (B, N) problem; allocated once, reused by every forward / graph replay."""

    def __init__(self, eng: "Engine", B: int, N: int, streams: int = 2):
        dev, t16 = eng.device, eng.torch16
        H, Mlp, T = eng.hidden, eng.mlp_dim, eng.num_tokens
        self.B, self.N, self.S, self.streams = B, N, T + N, streams
        n_seq = streams * B   # image blocks stacked as sequences: [ref | dist] or [ref | dist1 | dist2] (pairwise)
        rows, prow = n_seq * self.S, n_seq * N
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        self.patches16 = e(prow, eng.patch_elems, dt=t16)
        self.proj = e(prow, H)
        self.pos = e(prow, 2)
        self.scales = e(prow)
        self.x = e(rows, H)
        self.ln = e(rows, H, dt=t16)
        self.qkv = e(rows, 3 * H, dt=t16)
        self.att = e(rows, H, dt=t16)
        self.h1 = e(rows, Mlp, dt=t16)
        # per-row (sum, sum of squares) partials handed from a residual GEMM to the GEMM behind the LayerNorm
        self.ln_slots = max(int(_lib.load_library().vtq_gemm_ln_slots(H)), 1)
        self.stats = e(self.ln_slots, rows, 2)
        self.diff = e((streams - 1) * B, H)
        self.q = e((streams - 1) * B)
        self.tail_ws = torch.empty(max(eng.ctx.workspace_bytes((streams - 1) * B, H), 16), dtype=torch.uint8,
                                   device=dev)
        self.pos_idx = None    # optional int32 dumps for the parity tests
        self.scale_idx = None
        self.graph = None      # captured encoder+tail graph


class _WsView:
    """Pointers into a workspace's row-indexed buffers, offset to a sequence range (patch row p0, token row r0)."""

    def __init__(self, ws: _Workspace, p0: int, r0: int):
        off = lambda t, rows: None if t is None else C.c_void_p(t.data_ptr() + rows * t.stride(0) * t.element_size())
        self.patches16, self.proj, self.pos, self.scales = (off(ws.patches16, p0), off(ws.proj, p0), off(ws.pos, p0),
                                                            off(ws.scales, p0))
        self.x, self.ln, self.qkv, self.att, self.h1 = (off(ws.x, r0), off(ws.ln, r0), off(ws.qkv, r0), off(ws.att, r0),
                                                        off(ws.h1, r0))
        # statistics slots are [slot][row][2]: a row offset inside every slot (the kernels index slot * M + row with
        # M = the rows of THEIR launch, so split launches are only used without LayerNorm folding)
        self.stats = off(ws.stats.view(-1, 2), r0)
        self.pos_idx = off(ws.pos_idx, p0)
        self.scale_idx = off(ws.scale_idx, p0)


class Engine:
    """Packed weights + workspaces + launch sequence for one VTAMIQ module on one device."""

    def __init__(self, model, operand_dtype: str = "fp16", use_cuda_graph: bool = True,
                 prune_last_block: bool = True, fuse_layernorm: bool | None = None):
        if operand_dtype not in _DTYPES:
            raise ValueError(f"operand_dtype must be one of {list(_DTYPES)}")
        self.model = model
        self.operand_dtype = operand_dtype
        self.vtq16, self.torch16 = _DTYPES[operand_dtype]
        self.use_cuda_graph = use_cuda_graph
        self.prune_last_block = prune_last_block   # last block: quality-token row only after K/V (exact)
        # encoder LayerNorms carried by the GEMMs either side of them (needs >= 256 token rows).  Opt-in: measured
        # 3-4 % SLOWER per step at cfg2 than the separate LayerNorm kernel.  When the caller does
        # not choose, VTQ_FUSE_LN=1 turns it on (A/B runs).
        if fuse_layernorm is None:
            fuse_layernorm = os.environ.get("VTQ_FUSE_LN", "0") == "1"
        self.fuse_layernorm = bool(fuse_layernorm)
        self.device = None
        self.ctx = None
        self._sig = None
        self._ws: "OrderedDict[tuple[int, int, int], _Workspace]" = OrderedDict()
        self.max_workspaces = int(os.environ.get("VTQ_MAX_WORKSPACES", "4"))   # LRU bound on cached (B, N) workspaces
        self.static_weights = False   # True: skip the per-forward parameter version walk (weights promised frozen)
        self.dump_indices = False
        self.timeline = None   # when a list: (tag, start_event, end_event) per launch (bench.py roofline leg)
        # two kernel chains (halves of the sequence batch) on two streams inside the captured step: VTQ_SPLIT_STREAMS=1
        self.split_streams = os.environ.get("VTQ_SPLIT_STREAMS", "0") == "1"
        self._side_stream = None
        self.zigzag = os.environ.get("VTQ_ZIGZAG", "1") == "1"   # alternate the walk direction of consecutive kernels

    def _call(self, tag, name, *args):
        """ctx.call, optionally bracketed by CUDA events on the launching stream."""
        if self.timeline is None:
            self.ctx.call(name, *args)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.ctx.call(name, *args)
        e1.record()
        self.timeline.append((tag, e0, e1))

    # ------------------------------------------------------------------ weights
    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def _ensure_ready(self):
        p0 = next(self.model.parameters())
        if p0.device.type != "cuda":
            raise VtqError("vtamiq_b200: the model must live on a CUDA (sm_100) device; there is no CPU path. "
                           "Call model.to('cuda') first.")
        if self.device != p0.device:
            self.device = p0.device
            self.ctx = get_context(p0.device.index if p0.device.index is not None else torch.cuda.current_device())
            self._sig = None
            self._ws.clear()
        if self.static_weights and self._sig is not None:
            return
        sig = self._signature() + (self.fuse_layernorm,)
        if sig != self._sig:
            self._pack()
            self._sig = sig
            for ws in self._ws.values():
                ws.graph = None  # packed buffers were re-created: captured pointers are stale

    @torch.no_grad()
    def _pack(self):
        """(Re)build the kernel-side weight copies from the module's fp32 parameters."""
        m, t16 = self.model, self.torch16
        vit = m.transformer
        emb = vit.embeddings
        for p in m.parameters():
            if p.dtype != torch.float32:
                raise VtqError("vtamiq_b200 expects fp32 master parameters (model.to(device, torch.float32))")
        f32 = lambda t: t.detach().contiguous()
        h16 = lambda t: t.detach().to(t16).contiguous()
        self.hidden = vit.hidden_size
        self.heads = vit.config["num_heads"]
        self.mlp_dim = vit.config["mlp_dim"]
        self.num_tokens = emb.num_tokens
        self.patch = vit.config["patch_size"]
        self.patch_elems = 3 * self.patch * self.patch
        if self.hidden // self.heads != 64:
            raise VtqError("attention kernel supports head_dim 64 only")
        if self.patch_elems % 64 != 0:
            raise VtqError("patch size must give a K multiple of 64 for the embed GEMM")
        if emb.use_patch_embedding:
            self.w_pe = h16(emb.patch_embeddings.weight.reshape(self.hidden, -1))
            self.b_pe = f32(emb.patch_embeddings.bias)
        else:
            self.w_pe = self.b_pe = None
        self.cls = f32(emb.cls_token.reshape(-1)) if emb.use_cls_token else None
        self.extra = f32(emb.extra_tokens.reshape(-1, self.hidden)) if emb.use_extra_tokens else None
        self.n_extra = emb.num_extra_tokens
        if emb.use_pos_embedding:
            self.pos_table = f32(emb.positional_embeddings.positional_embeddings[0])
            self.pos_grid = emb.positional_embeddings.width_pos_embeddings
        else:
            self.pos_table, self.pos_grid = None, 0
        if emb.use_scale_embedding:
            self.scale_table = f32(emb.scale_embeddings.scale_embeddings[0])
            self.num_scales = emb.scale_embeddings.num_scales
        else:
            self.scale_table, self.num_scales = None, 0
        self.layers = []

        fold = lambda w, b, ln_w, ln_b: fold_layernorm(w, b, ln_w, ln_b, t16)

        for L in vit.encoder.layers:
            a = L.attn
            folded = {}
            if self.fuse_layernorm:
                w_qkv32 = torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)
                b_qkv32 = torch.cat([a.query.bias, a.key.bias, a.value.bias], 0)
                w_qkv_f, b_qkv_f, cs_qkv = fold(w_qkv32, b_qkv32, L.attention_norm.weight, L.attention_norm.bias)
                w_fc1_f, b_fc1_f, cs_fc1 = fold(L.ffn.fc1.weight, L.ffn.fc1.bias, L.ffn_norm.weight, L.ffn_norm.bias)
                folded = dict(w_qkv_f=w_qkv_f, b_qkv_f=b_qkv_f, cs_qkv=cs_qkv, w_fc1_f=w_fc1_f, b_fc1_f=b_fc1_f,
                              cs_fc1=cs_fc1)
            self.layers.append(_LayerPack(
                **folded,
                ln1_w=f32(L.attention_norm.weight), ln1_b=f32(L.attention_norm.bias),
                w_qkv=h16(torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)),
                b_qkv=f32(torch.cat([a.query.bias, a.key.bias, a.value.bias], 0)),
                w_o=h16(a.out.weight), b_o=f32(a.out.bias),
                g1=f32(L.ls1.gamma) if vit.use_layer_scale else None,
                ln2_w=f32(L.ffn_norm.weight), ln2_b=f32(L.ffn_norm.bias),
                w_fc1=h16(L.ffn.fc1.weight), b_fc1=f32(L.ffn.fc1.bias),
                w_fc2=h16(L.ffn.fc2.weight), b_fc2=f32(L.ffn.fc2.bias),
                g2=f32(L.ls2.gamma) if vit.use_layer_scale else None,
                # Houlsby adapters: the VTAMIQ path uses adapter pair 0 of a layer (backbone.py:54-59)
                **(dict(ad1=self._pack_adapter(L.adapters[0][0], h16, f32), ad2=self._pack_adapter(L.adapters[0][1], h16, f32))
                   if getattr(L, "use_adapters", False) else dict(ad1=None, ad2=None)),
            ))
        self.use_adapters = bool(getattr(vit, "use_adapters", False))
        self.ln_eps = float(vit.encoder.encoder_norm.eps)
        self.lnf_w, self.lnf_b = f32(vit.encoder.encoder_norm.weight), f32(vit.encoder.encoder_norm.bias)
        self.diff_gamma = f32(m.diff_scale.gamma) if hasattr(m.diff_scale, "gamma") else None
        # DiffNet + head parameter list in the order vtq_diffnet_head documents
        plist = []
        groups = [mod for mod in m.quality_decoder if hasattr(mod, "body")]
        self.num_rgs = len(groups)
        self.num_rcabs = 0
        self.ca_hidden = 32
        for grp in groups:
            rcabs = list(grp.body)[:-1]
            self.num_rcabs = len(rcabs)
            for rc in rcabs:
                prelu, conv, ca = rc.body[1], rc.body[2], rc.body[4]
                if prelu.weight.numel() != 1:
                    raise VtqError("DiffNet PReLU must have a single slope")
                down, up = ca.conv_du[1], ca.conv_du[4]
                self.ca_hidden = down.weight.shape[0]
                plist += [f32(prelu.weight), f32(conv.weight.squeeze(-1)), f32(conv.bias),
                          f32(down.weight.squeeze(-1)), f32(down.bias), f32(up.weight.squeeze(-1)), f32(up.bias)]
            gconv = grp.body[-1]
            plist += [f32(gconv.weight.squeeze(-1)), f32(gconv.bias)]
        if self.num_rgs > 0:
            fconv = m.quality_decoder[-1]
            plist += [f32(fconv.weight.squeeze(-1)), f32(fconv.bias)]
        else:
            plist += [None, None]
        lin1, pre, lin2 = m.q_predictor[1], m.q_predictor[2], m.q_predictor[4]
        self.head_hidden = lin1.weight.shape[0]
        plist += [f32(lin1.weight), f32(lin1.bias), f32(pre.weight), f32(lin2.weight), f32(lin2.bias)]
        self._tail_tensors = plist  # keep alive (detached views of the live parameters)
        arr = (C.c_void_p * len(plist))(*[None if t is None else t.data_ptr() for t in plist])
        self._tail_params = arr
        self.token_num = int(getattr(m, "token_num", 0))

    # ------------------------------------------------------------------ workspaces
    def workspace(self, B: int, N: int, streams: int = 2) -> _Workspace:
        self._ensure_ready()
        key = (B, N, streams)
        ws = self._ws.get(key)
        if ws is None:
            while len(self._ws) >= max(self.max_workspaces, 1):   # least-recently-used (B, N) goes first
                self._ws.popitem(last=False)
            with torch.cuda.device(self.device):
                ws = self._ws[key] = _Workspace(self, B, N, streams)
        else:
            self._ws.move_to_end(key)
        return ws

    def clear_workspaces(self):
        """Drop every cached activation workspace and captured graph (they are re-created on demand)."""
        self._ws.clear()

    # ------------------------------------------------------------------ launch sequence
    @staticmethod
    def _pack_adapter(ad, h16, f32):
        l1, l2 = ad.adapter[0], ad.adapter[2]
        if l1.weight.shape[0] % 64 or l1.weight.shape[1] % 64:
            raise VtqError("adapter widths must be multiples of 64 for the GEMM kernels")
        return (h16(l1.weight), f32(l1.bias), h16(l2.weight), f32(l2.bias), int(l1.weight.shape[0]))

    def _adapter(self, c, w, ad, src16, lds, W, bias, rows, K, gamma, dt, st):
        """x += gamma * adapter(h) with h = src16 W^T + bias: h is produced once more as a 16-bit matrix (the GEMM
        operand precision of this engine), then the two adapter GEMMs; the un-adapted branch gamma * h itself was
        added by the caller's residual GEMM (transformer.py:277-279, :282-284: x + ls(h + adapter(h)))."""
        w1, b1, w2, b2, width = ad
        H = self.hidden
        c("gemm_adapter_in", "vtq_gemm", src16, lds, W, bias, rows, H, K, dt, EPI_BIAS_H, w.ln, 0, None, st)
        c("gemm_adapter_1", "vtq_gemm", w.ln, 0, _ptr(w1), _ptr(b1), rows, width, H, dt, EPI_BIAS_GELU_H, w.att, 0,
          None, st)
        c("gemm_adapter_2", "vtq_gemm", w.att, 0, _ptr(w2), _ptr(b2), rows, H, width, dt, EPI_BIAS_RESID_F32, w.x, 0,
          gamma, st)

    def _encode_part(self, ws: _Workspace, embedded: bool, seq0: int, n_seq: int, st):
        """Patch projection, embedding and the encoder blocks for sequences [seq0, seq0 + n_seq) on stream ``st``."""
        c, dt = self._call, self.vtq16
        N, S, H = ws.N, ws.S, self.hidden
        rows, prow = n_seq * S, n_seq * N
        w = _WsView(ws, seq0 * N, seq0 * S)
        # Zig-zag traversal: consecutive row-/tile-walking kernels of the chain alternate their walk direction
        # (vtq_set_reverse), so each one starts on the rows its predecessor wrote last — the part of the 0.65 GB of
        # per-block activations that is still in the 126 MB L2.  Results are identical either way.
        rev = [True]   # the embedding wrote x in increasing row order: the first LayerNorm starts from the end

        def d(tag, name, *args):
            if self.zigzag:
                self.ctx.call("vtq_set_reverse", int(rev[0]))
                rev[0] = not rev[0]
            c(tag, name, *args)
        if not embedded:
            c("gemm_embed", "vtq_gemm", w.patches16, 0, _ptr(self.w_pe), _ptr(self.b_pe), prow, H,
              self.patch_elems, dt, EPI_BIAS_F32, w.proj, 0, None, st)
        c("embed_assemble", "vtq_embed_assemble", w.proj, w.pos,
          w.scales if self.scale_table is not None else None,
          _ptr(self.pos_table), self.pos_grid, _ptr(self.scale_table), self.num_scales, _ptr(self.cls),
          _ptr(self.extra), self.n_extra, n_seq, N, H, w.x,
          w.pos_idx if self.dump_indices else None, w.scale_idx if self.dump_indices else None, st)
        eps = self.ln_eps
        n_layers = len(self.layers)
        # Folded LayerNorms: ws.ln holds the RAW 16-bit copy of x, ws.stats each row's (sum, sum of squares) partials;
        # the residual GEMMs (out-projection, fc2) refresh both, the GEMMs behind a LayerNorm (QKV, fc1) consume them.
        fold = self.fuse_layernorm and rows >= 256 and not self.use_adapters
        slots_in = 1
        if fold:
            c("rowstats_cast", "vtq_rowstats_cast", w.x, rows, H, w.ln, w.stats, dt, st)
        for li, L in enumerate(self.layers):
            if fold:
                d("gemm_qkv", "vtq_gemm_ln", w.ln, 0, _ptr(L.w_qkv_f), _ptr(L.b_qkv_f), rows, 3 * H, H, dt,
                  EPI_BIAS_H, w.qkv, 0, None, w.stats, slots_in, _ptr(L.cs_qkv), eps, None, None, st)
            else:
                d("layernorm", "vtq_layernorm", w.x, 0, _ptr(L.ln1_w), _ptr(L.ln1_b), eps, rows, H, w.ln,
                  dt, st)
                d("gemm_qkv", "vtq_gemm", w.ln, 0, _ptr(L.w_qkv), _ptr(L.b_qkv), rows, 3 * H, H, dt, EPI_BIAS_H,
                  w.qkv, 0, None, st)
            if self.prune_last_block and li == n_layers - 1 and L.ad1 is None:
                # Only the quality token of each sequence survives the encoder (transformer.py:634, vtamiq.py:104-108):
                # in the last block K/V still need every row, but attention output, out-projection, LayerNorm and the
                # MLP are evaluated for that one row per sequence (row stride S*H picks it out of x / att).
                tok = self.token_num
                c("attention_tok", "vtq_attention_fwd", w.qkv, w.att, n_seq, S, self.heads, dt,
                  tok + 1, st)
                x_tok = C.c_void_p(w.x.value + tok * H * 4)
                att_tok = C.c_void_p(w.att.value + tok * H * 2)
                c("gemm_out_tok", "vtq_gemm", att_tok, S * H, _ptr(L.w_o), _ptr(L.b_o), n_seq, H, H, dt,
                  EPI_BIAS_RESID_F32, x_tok, S * H, _ptr(L.g1), st)
                c("layernorm_tok", "vtq_layernorm", x_tok, S * H, _ptr(L.ln2_w), _ptr(L.ln2_b), eps, n_seq, H,
                  w.ln, dt, st)
                c("gemm_fc1_tok", "vtq_gemm", w.ln, 0, _ptr(L.w_fc1), _ptr(L.b_fc1), n_seq, self.mlp_dim, H, dt,
                  EPI_BIAS_GELU_H, w.h1, 0, None, st)
                c("gemm_fc2_tok", "vtq_gemm", w.h1, 0, _ptr(L.w_fc2), _ptr(L.b_fc2), n_seq, H, self.mlp_dim, dt,
                  EPI_BIAS_RESID_F32, x_tok, S * H, _ptr(L.g2), st)
                continue
            d("attention", "vtq_attention_fwd", w.qkv, w.att, n_seq, S, self.heads, dt, 0, st)
            if fold:
                d("gemm_out", "vtq_gemm_ln", w.att, 0, _ptr(L.w_o), _ptr(L.b_o), rows, H, H, dt,
                  EPI_BIAS_RESID_F32, w.x, 0, _ptr(L.g1), None, 0, None, 0.0, w.ln, w.stats, st)
                slots_in = ws.ln_slots
                d("gemm_fc1", "vtq_gemm_ln", w.ln, 0, _ptr(L.w_fc1_f), _ptr(L.b_fc1_f), rows, self.mlp_dim, H,
                  dt, EPI_BIAS_GELU_H, w.h1, 0, None, w.stats, slots_in, _ptr(L.cs_fc1), eps, None,
                  None, st)
                d("gemm_fc2", "vtq_gemm_ln", w.h1, 0, _ptr(L.w_fc2), _ptr(L.b_fc2), rows, H, self.mlp_dim, dt,
                  EPI_BIAS_RESID_F32, w.x, 0, _ptr(L.g2), None, 0, None, 0.0, w.ln, w.stats, st)
                continue
            d("gemm_out", "vtq_gemm", w.att, 0, _ptr(L.w_o), _ptr(L.b_o), rows, H, H, dt, EPI_BIAS_RESID_F32,
              w.x, 0, _ptr(L.g1), st)
            if L.ad1 is not None:
                self._adapter(c, w, L.ad1, w.att, 0, _ptr(L.w_o), _ptr(L.b_o), rows, H, _ptr(L.g1), dt, st)
            d("layernorm", "vtq_layernorm", w.x, 0, _ptr(L.ln2_w), _ptr(L.ln2_b), eps, rows, H, w.ln, dt, st)
            d("gemm_fc1", "vtq_gemm", w.ln, 0, _ptr(L.w_fc1), _ptr(L.b_fc1), rows, self.mlp_dim, H, dt,
              EPI_BIAS_GELU_H, w.h1, 0, None, st)
            d("gemm_fc2", "vtq_gemm", w.h1, 0, _ptr(L.w_fc2), _ptr(L.b_fc2), rows, H, self.mlp_dim, dt,
              EPI_BIAS_RESID_F32, w.x, 0, _ptr(L.g2), st)
            if L.ad2 is not None:
                self._adapter(c, w, L.ad2, w.h1, 0, _ptr(L.w_fc2), _ptr(L.b_fc2), rows, self.mlp_dim, _ptr(L.g2), dt, st)
        if self.zigzag:
            self.ctx.call("vtq_set_reverse", 0)

    def _encode_and_score(self, ws: _Workspace, embedded: bool, tail: bool = True):
        """Everything after the inputs sit in ws.patches16 / ws.proj, ws.pos, ws.scales.  ``tail=False`` stops after
        the encoder with ws.diff = LN(cls_ref) - LN(cls_dist) (no diff_scale): the differentiable tail takes over.

        Sequences are independent until the CLS difference, so with ``split_streams`` the two halves of the sequence
        batch are encoded as two kernel chains on two streams (joined before cls_diff): the memory-bound kernels of
        one chain (LayerNorm, embedding) can then share the SMs with the tensor-bound kernels of the other."""
        c, st = self._call, _stream(self.device)
        B, S, H = ws.B, ws.S, self.hidden
        n_seq = ws.streams * B
        eps = self.ln_eps
        if self.dump_indices and ws.pos_idx is None:   # (before the buffer views of _encode_part are taken)
            ws.pos_idx = torch.zeros(n_seq * ws.N, dtype=torch.int32, device=self.device)
            ws.scale_idx = torch.zeros(n_seq * ws.N, dtype=torch.int32, device=self.device)
        split = (self.split_streams and self.timeline is None and not self.dump_indices and not self.fuse_layernorm
                 and n_seq >= 2 and (n_seq // 2) * S >= 256)
        if split:
            main = torch.cuda.current_stream(self.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            side = self._side_stream
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
            h0 = n_seq // 2
            with torch.cuda.stream(side):
                self._encode_part(ws, embedded, h0, n_seq - h0, C.c_void_p(side.cuda_stream))
                join.record(side)
            self._encode_part(ws, embedded, 0, h0, st)
            main.wait_event(join)
        else:
            self._encode_part(ws, embedded, 0, n_seq, st)
        for k in range(1, ws.streams):   # every distorted block against the (once-encoded) reference block
            c("cls_diff", "vtq_cls_diff", _ptr(ws.x), C.c_void_p(ws.x.data_ptr() + k * B * S * H * 4), B, S, H,
              self.token_num, _ptr(self.lnf_w), _ptr(self.lnf_b), eps, _ptr(self.diff_gamma) if tail else None,
              C.c_void_p(ws.diff.data_ptr() + (k - 1) * B * H * 4), st)
        if not tail:
            return
        nq = (ws.streams - 1) * B
        c("diffnet_head", "vtq_diffnet_head", _ptr(ws.diff), self._tail_params, len(self._tail_params), self.num_rgs,
          self.num_rcabs, H, self.ca_hidden, self.head_hidden, nq, _ptr(ws.q), _ptr(ws.tail_ws), st)

    def run(self, ws: _Workspace, embedded: bool = False, tail: bool = True):
        """Encode + score the staged inputs; uses a captured CUDA graph per workspace when enabled."""
        with torch.cuda.device(self.device):
            if not self.use_cuda_graph or self.dump_indices or self.timeline is not None:
                self._encode_and_score(ws, embedded, tail)
                return
            key = ("emb" if embedded else "patch", self.prune_last_block, self.fuse_layernorm, tail, self.split_streams, self.zigzag)
            if ws.graph is None or ws.graph[0] != key:
                # warm-up outside capture (cudaFuncSetAttribute, lazy module load), then capture
                self._encode_and_score(ws, embedded, tail)
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._encode_and_score(ws, embedded, tail)
                ws.graph = (key, g)
            ws.graph[1].replay()

    # ------------------------------------------------------------------ input staging
    def stage_patches(self, ws: _Workspace, patches, pos, scales):
        """Reference-style inputs: tuples (ref, dist) of (B,N,3,P,P) fp32 [or (B,N,H) embedded], (B,N,2), (B,N)."""
        c, st = self.ctx.call, _stream(self.device)
        B, N = ws.B, ws.N
        embedded = patches[0].dim() == 3
        dev = self.device
        # the reference runs on whatever device its inputs live on; here raw pointers go to the kernels, so inputs are
        # brought to the model's device first (a CPU tensor's data_ptr must never reach a launch)
        on_dev = lambda t: t if t.device == dev else t.to(dev, non_blocking=True)
        for img in range(ws.streams):
            p = on_dev(patches[img])
            if p.dtype != torch.float32 or not p.is_contiguous():
                p = p.to(torch.float32).contiguous()
            if embedded:
                ws.proj[img * B * N:(img + 1) * B * N].copy_(p.reshape(B * N, self.hidden))
            else:
                c("vtq_cast_rows", _ptr(p), C.c_void_p(ws.patches16[img * B * N].data_ptr()), p.numel(), self.vtq16, st)
            if self.pos_table is not None:
                ws.pos[img * B * N:(img + 1) * B * N].copy_(on_dev(pos[img]).reshape(B * N, 2))
            if self.scale_table is not None:
                if scales is None or scales[img] is None:
                    raise ValueError("Model uses scale embedding but scales is passed as None.")
                ws.scales[img * B * N:(img + 1) * B * N].copy_(on_dev(scales[img]).reshape(B * N))
        return embedded
