"""Differentiable tail: DiffNet + quality head with a hand-written backward (vtq_tail_train_fwd / vtq_tail_bwd).

What the reference does here is plain ``torch.autograd`` over ``diff_scale`` → ``quality_decoder`` → ``q_predictor``
(modules/vtamiq/vtamiq.py:111-117) inside ``train.py:317-322``.  With the encoder frozen (``set_freeze_state``,
backbone.py:62-106) those are the only parameters that train, so the encoder can stay on the forward-only kernels
and only this tail needs gradients.  Training-mode semantics kept: ``DropPath`` on every ResidualGroup branch
(channel_attention.py:26-29; ``rg_path_drop``, per-pair Bernoulli mask / keep_prob drawn with torch's generator in the
reference's order); ``predictor_dropout`` > 0 in training mode is not supported (raises).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import _ptr, _stream


def tail_parameters(model):
    """The live ``nn.Parameter`` objects in the order vtq_diffnet_head / vtq_tail_bwd list them (None = absent)."""
    plist = []
    groups = [mod for mod in model.quality_decoder if hasattr(mod, "body")]
    for grp in groups:
        for rc in list(grp.body)[:-1]:
            prelu, conv, ca = rc.body[1], rc.body[2], rc.body[4]
            down, up = ca.conv_du[1], ca.conv_du[4]
            plist += [prelu.weight, conv.weight, conv.bias, down.weight, down.bias, up.weight, up.bias]
        gconv = grp.body[-1]
        plist += [gconv.weight, gconv.bias]
    if groups:
        fconv = model.quality_decoder[-1]
        plist += [fconv.weight, fconv.bias]
    else:
        plist += [None, None]
    lin1, pre, lin2 = model.q_predictor[1], model.q_predictor[2], model.q_predictor[4]
    plist += [lin1.weight, lin1.bias, pre.weight, lin2.weight, lin2.bias]
    return plist


class _Tail(torch.autograd.Function):
    """q = head(DiffNet(gamma * d0)); inputs after the fixed ones: gamma (or None), then the tail parameters."""

    @staticmethod
    def forward(ctx, eng, dims, d0, drop_scale, gamma, *params):
        num_rgs, num_rcabs, H, ca, hh = dims
        B = d0.shape[0]
        dev = d0.device
        lib = _lib.load_library()
        n_saved = int(lib.vtq_tail_saved_floats(B, num_rgs, num_rcabs, H, ca, hh))
        saved = torch.empty(n_saved, dtype=torch.float32, device=dev)
        q = torch.empty(B, dtype=torch.float32, device=dev)
        counters = torch.empty(32768, dtype=torch.uint8, device=dev)
        dets = [None if p is None else p.detach().contiguous() for p in params]
        arr = (C.c_void_p * len(dets))(*[None if t is None else t.data_ptr() for t in dets])
        g = None if gamma is None else gamma.detach().contiguous()
        with torch.cuda.device(dev):
            eng.ctx.call("vtq_tail_train_fwd", _ptr(d0), _ptr(g), arr, len(dets), num_rgs, num_rcabs, H, ca, hh, B,
                         _ptr(drop_scale), _ptr(saved), _ptr(q), _ptr(counters), _stream(dev))
        ctx.eng, ctx.dims, ctx.n_params = eng, dims, len(dets)
        ctx.has_gamma = gamma is not None
        ctx.save_for_backward(d0, saved, *( [] if drop_scale is None else [drop_scale]), *([g] if g is not None else []),
                              *[t for t in dets if t is not None])
        ctx.has_drop = drop_scale is not None
        ctx.present = [t is not None for t in dets]
        return q

    @staticmethod
    def backward(ctx, dq):
        eng = ctx.eng
        num_rgs, num_rcabs, H, ca, hh = ctx.dims
        tensors = list(ctx.saved_tensors)
        d0, saved = tensors[0], tensors[1]
        k = 2
        drop_scale = None
        if ctx.has_drop:
            drop_scale = tensors[k]
            k += 1
        gamma = None
        if ctx.has_gamma:
            gamma = tensors[k]
            k += 1
        dets = []
        for present in ctx.present:
            dets.append(tensors[k] if present else None)
            k += present
        B, dev = d0.shape[0], d0.device
        lib = _lib.load_library()
        need = ctx.needs_input_grad   # (eng, dims, d0, drop_scale, gamma, *params)
        grads = [torch.empty_like(t) if (t is not None and need[5 + i]) else None for i, t in enumerate(dets)]
        dgamma = torch.empty_like(gamma) if (gamma is not None and need[4]) else None
        d_d0 = torch.empty_like(d0) if need[2] else None
        ws = torch.empty(int(lib.vtq_tail_bwd_workspace_bytes(B, H)), dtype=torch.uint8, device=dev)
        parr = (C.c_void_p * len(dets))(*[None if t is None else t.data_ptr() for t in dets])
        garr = (C.c_void_p * len(dets))(*[None if t is None else t.data_ptr() for t in grads])
        dq = dq.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            eng.ctx.call("vtq_tail_bwd", _ptr(dq), _ptr(d0), _ptr(gamma), parr, garr, len(dets), num_rgs, num_rcabs, H,
                         ca, hh, B, _ptr(drop_scale), _ptr(saved), _ptr(dgamma), _ptr(d_d0), _ptr(ws), _stream(dev))
        return (None, None, d_d0, None, dgamma, *grads)


def draw_drop_scale(model, B, n_groups, device):
    """Per-pair DropPath factors of the ResidualGroups, [n_groups][B], drawn the way timm's DropPath draws them
    (``x.new_empty(B, 1, 1).bernoulli_(keep).div_(keep)``, one draw per group in forward order) so that, for the same
    generator state, the masks are the ones the reference would use.  None when DropPath is inactive."""
    p = float(getattr(model, "rg_path_drop", 0.0))
    override = getattr(model, "_drop_scale_override", None)
    if override is not None:
        return override.to(device=device, dtype=torch.float32).contiguous()
    if not model.training or p == 0.0 or n_groups == 0:
        return None
    keep = 1.0 - p
    rows = []
    for _ in range(n_groups):
        m = torch.empty(B, 1, 1, dtype=torch.float32, device=device).bernoulli_(keep)
        if keep > 0.0:
            m.div_(keep)
        rows.append(m.view(B))
    return torch.stack(rows).contiguous()


def run_tail(model, ws, n_dist):
    """Scores of the ``n_dist`` distorted blocks staged in ``ws`` (ws.diff holds the unscaled differences)."""
    eng = model.engine
    if model.training and float(model.predictor_dropout) > 0.0:
        raise NotImplementedError("vtamiq_b200: predictor_dropout > 0 in training mode is not supported")
    B = ws.B
    dims = (eng.num_rgs, eng.num_rcabs, eng.hidden, eng.ca_hidden, eng.head_hidden)
    params = tail_parameters(model)
    gamma = model.diff_scale.gamma if hasattr(model.diff_scale, "gamma") else None
    out = []
    for k in range(n_dist):
        d0 = ws.diff[k * B:(k + 1) * B].clone()     # the workspace is reused by the next forward
        drop = draw_drop_scale(model, B, eng.num_rgs, d0.device)
        out.append(_Tail.apply(eng, dims, d0, drop, gamma, *params))
    return tuple(out)
