"""Per-kernel parity on a real B200, through the C-ABI (ctypes), against torch fp32 math / the oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import synth  # noqa: E402
from oracle import patch_oracle, vtamiq_oracle  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from vtamiq_b200 import _lib
    assert torch.cuda.is_available(), "these tests need the GPU box"
    return _lib.get_context(0)


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def ST():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


DT = {"fp16": (0, torch.float16), "bf16": (1, torch.bfloat16)}


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 768, 768), (1000, 2304, 768), (257, 3072, 768),
                                   (515, 768, 3072), (32064, 768, 768)])
def test_gemm_bias_h(ctx, dt, M, N, K):
    code, tdt = DT[dt]
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(tdt)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(tdt)
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=tdt)
    ctx.call("vtq_gemm", P(A), 0, P(W), P(b), M, N, K, code, 0, P(out), 0, None, ST())
    torch.cuda.synchronize()
    want = A.float() @ W.float().t() + b
    err = (out.float() - want).abs().max().item()
    tol = 4e-2 if dt == "bf16" else 4e-3   # output rounding: |out| up to ~6, eps 2^-8 / 2^-11
    assert torch.isfinite(out.float()).all()
    assert err < tol, err


@pytest.mark.parametrize("dt", ["fp16"])
def test_gemm_gelu(ctx, dt):
    code, tdt = DT[dt]
    M, N, K = 700, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).to(tdt)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(tdt)
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=tdt)
    ctx.call("vtq_gemm", P(A), 0, P(W), P(b), M, N, K, code, 1, P(out), 0, None, ST())
    torch.cuda.synchronize()
    want = torch.nn.functional.gelu(A.float() @ W.float().t() + b)
    assert (out.float() - want).abs().max().item() < 4e-3


@pytest.mark.parametrize("M,N,K,use_gamma", [(300, 768, 768, False), (1000, 768, 3072, True), (64, 768, 768, True)])
def test_gemm_f32_store_and_residual(ctx, M, N, K, use_gamma):
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    b = torch.randn(N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5 if use_gamma else None
    base = A.float() @ W.float().t() + b
    out = torch.full((M, N), float("nan"), device="cuda")
    ctx.call("vtq_gemm", P(A), 0, P(W), P(b), M, N, K, 0, 2, P(out), 0, None, ST())
    torch.cuda.synchronize()
    assert (out - base).abs().max().item() < 2e-3
    x0 = torch.randn(M, N, device="cuda", generator=g)
    x = x0.clone()
    ctx.call("vtq_gemm", P(A), 0, P(W), P(b), M, N, K, 0, 3, P(x), 0, P(gamma), ST())
    torch.cuda.synchronize()
    want = x0 + (base * gamma if use_gamma else base)
    assert (x - want).abs().max().item() < 2e-3


def test_gemm_rejects_bad_shapes(ctx):
    from vtamiq_b200._lib import VtqError
    A = torch.zeros(128, 100, device="cuda", dtype=torch.float16)
    W = torch.zeros(128, 100, device="cuda", dtype=torch.float16)
    b = torch.zeros(128, device="cuda")
    o = torch.zeros(128, 128, device="cuda", dtype=torch.float16)
    with pytest.raises(VtqError, match="multiple of 64"):
        ctx.call("vtq_gemm", P(A), 0, P(W), P(b), 128, 128, 100, 0, 0, P(o), 0, None, ST())


# ------------------------------------------------------------------------------------------ GEMM + folded LayerNorm
def _ln_slots(N):
    from vtamiq_b200 import _lib
    return int(_lib.load_library().vtq_gemm_ln_slots(N))


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("M,N,K,use_gamma", [(300, 768, 768, False), (1000, 768, 3072, True), (257, 1024, 1024, True),
                                             (32064, 768, 768, False)])
def test_gemm_ln_produce(ctx, dt, M, N, K, use_gamma):
    """Residual epilogue that also emits the raw 16-bit rows and the per-row (sum, sum of squares) partials."""
    code, tdt = DT[dt]
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g).to(tdt)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(tdt)
    b = torch.randn(N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5 if use_gamma else None
    x0 = torch.randn(M, N, device="cuda", generator=g) * 2 + 0.3
    x = x0.clone()
    slots = _ln_slots(N)
    raw = torch.full((M, N), float("nan"), device="cuda", dtype=tdt)
    stats = torch.full((slots, M, 2), float("nan"), device="cuda")
    ctx.call("vtq_gemm_ln", P(A), 0, P(W), P(b), M, N, K, code, 3, P(x), 0, P(gamma), None, 0, None, 0.0, P(raw),
             P(stats), ST())
    torch.cuda.synchronize()
    base = A.float() @ W.float().t() + b
    want = x0 + (base * gamma if use_gamma else base)
    tol = 2e-2 if dt == "bf16" else 2e-3
    assert (x - want).abs().max().item() < tol
    assert torch.equal(raw, x.to(tdt))                       # the 16-bit copy is the rounded fp32 row, bit for bit
    tot = stats.double().sum(0)
    assert torch.allclose(tot[:, 0], x.double().sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(tot[:, 1], (x.double() ** 2).sum(1), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("M,N,K,gelu,slots", [(300, 2304, 768, False, 1), (1000, 3072, 768, True, 8),
                                              (32064, 2304, 768, False, 8), (515, 4096, 1024, True, 16)])
def test_gemm_ln_consume(ctx, dt, M, N, K, gelu, slots):
    """out = epi(LN(x) W^T + b) computed from RAW 16-bit rows, folded weights and per-row statistics."""
    code, tdt = DT[dt]
    g = torch.Generator(device="cuda").manual_seed(N + K + slots)
    x = torch.randn(M, K, device="cuda", generator=g) * 1.7 + 0.4
    x[:, 5] *= 12.0                                           # an outlier channel, as real ViT streams have
    ln_w = torch.rand(K, device="cuda", generator=g) + 0.5
    ln_b = torch.randn(K, device="cuda", generator=g) * 0.2
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    eps = 1e-6
    # fold exactly as Engine._pack does
    Wf = (W * ln_w[None, :]).to(tdt)
    bf = b + W @ ln_b
    cs = Wf.float().sum(1).contiguous()
    # statistics split over `slots` column groups (what the producing GEMM leaves behind)
    stats = torch.zeros(slots, M, 2, device="cuda")
    for s_, cols in enumerate(torch.arange(K, device="cuda").chunk(slots)):
        stats[s_, :, 0] = x[:, cols].sum(1)
        stats[s_, :, 1] = (x[:, cols] ** 2).sum(1)
    raw = x.to(tdt)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=tdt)
    ctx.call("vtq_gemm_ln", P(raw), 0, P(Wf), P(bf), M, N, K, code, 1 if gelu else 0, P(out), 0, None, P(stats), slots,
             P(cs), eps, None, None, ST())
    torch.cuda.synchronize()
    want = torch.nn.functional.layer_norm(x, (K,), ln_w, ln_b, eps) @ W.t() + b
    if gelu:
        want = torch.nn.functional.gelu(want)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - want).abs().max().item()
    assert err < (8e-2 if dt == "bf16" else 8e-3), err


def test_rowstats_cast(ctx):
    for hidden in (768, 1024):
        x = torch.randn(1003, hidden, device="cuda") * 3 + 0.7
        raw = torch.empty(1003, hidden, device="cuda", dtype=torch.float16)
        stats = torch.empty(1, 1003, 2, device="cuda")
        ctx.call("vtq_rowstats_cast", P(x), 1003, hidden, P(raw), P(stats), 0, ST())
        torch.cuda.synchronize()
        assert torch.equal(raw, x.half())
        assert torch.allclose(stats[0, :, 0], x.sum(1), rtol=1e-5, atol=1e-3)
        assert torch.allclose(stats[0, :, 1], (x * x).sum(1), rtol=1e-5, atol=1e-2)


def test_gemm_ln_rejects_misuse(ctx):
    from vtamiq_b200._lib import VtqError
    A = torch.zeros(128, 768, device="cuda", dtype=torch.float16)
    W = torch.zeros(768, 768, device="cuda", dtype=torch.float16)
    b = torch.zeros(768, device="cuda")
    x = torch.zeros(128, 768, device="cuda")
    raw = torch.zeros(128, 768, device="cuda", dtype=torch.float16)
    st = torch.zeros(8, 128, 2, device="cuda")
    with pytest.raises(VtqError, match="M >= 256"):
        ctx.call("vtq_gemm_ln", P(A), 0, P(W), P(b), 128, 768, 768, 0, 3, P(x), 0, None, None, 0, None, 0.0, P(raw),
                 P(st), ST())
    A = torch.zeros(512, 768, device="cuda", dtype=torch.float16)
    x = torch.zeros(512, 768, device="cuda")
    with pytest.raises(VtqError, match="exactly one"):
        ctx.call("vtq_gemm_ln", P(A), 0, P(W), P(b), 512, 768, 768, 0, 3, P(x), 0, None, None, 0, None, 0.0, None,
                 None, ST())


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, n_seq, S, heads):
    H = heads * 64
    x = qkv.float().view(n_seq, S, 3, heads, 64)
    q, k, v = x[:, :, 0].permute(0, 2, 1, 3), x[:, :, 1].permute(0, 2, 1, 3), x[:, :, 2].permute(0, 2, 1, 3)
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(n_seq * S, H)


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("n_seq,S,heads", [(2, 128, 2), (3, 65, 12), (2, 129, 12), (4, 501, 12), (2, 257, 12),
                                           (1, 1, 12), (1, 1300, 4)])
def test_attention(ctx, dt, n_seq, S, heads):
    code, tdt = DT[dt]
    H = heads * 64
    g = torch.Generator(device="cuda").manual_seed(S)
    qkv = (torch.randn(n_seq * S, 3 * H, device="cuda", generator=g) * 1.5).to(tdt)
    out = torch.full((n_seq * S, H), float("nan"), device="cuda", dtype=tdt)
    ctx.call("vtq_attention_fwd", P(qkv), P(out), n_seq, S, heads, code, 0, ST())
    torch.cuda.synchronize()
    want = _attn_ref(qkv, n_seq, S, heads)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - want).abs().max().item()
    assert err < (3e-2 if dt == "bf16" else 4e-3), err


def test_attention_long_sequence_5001(ctx):
    """BASELINE cfg4's sequence length as a unit test: S = 5001 (40 key tiles per query row, ragged last tile)."""
    n_seq, S, heads = 1, 5001, 2
    H = heads * 64
    g = torch.Generator(device="cuda").manual_seed(5001)
    qkv = (torch.randn(n_seq * S, 3 * H, device="cuda", generator=g) * 1.5).half()
    out = torch.full((n_seq * S, H), float("nan"), device="cuda", dtype=torch.float16)
    ctx.call("vtq_attention_fwd", P(qkv), P(out), n_seq, S, heads, 0, 0, ST())
    torch.cuda.synchronize()
    want = _attn_ref(qkv, n_seq, S, heads)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - want).abs().max().item() < 4e-3


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
def test_attention_rescale_is_forced_by_score_spikes(ctx, dt):
    """The online softmax keeps its reference maximum until the true maximum has grown by more than 2^8 (lazy
    rescale); N(0, 1.5) inputs never get there.  Here every query row meets keys whose scores climb by ~40 (raw
    score units, i.e. ~2^7 .. 2^58 in the exp2 domain) from one key tile to the next: key tile j carries one key
    aligned with the queries' common direction, scaled by j.  Every tile boundary then forces the O / l correction,
    and the planted keys dominate the softmax — the value rows they select must come out."""
    code, tdt = DT[dt]
    n_seq, S, heads = 2, 900, 3
    H = heads * 64
    g = torch.Generator(device="cuda").manual_seed(77)
    qkv = (torch.randn(n_seq, S, 3, heads, 64, device="cuda", generator=g) * 0.5)
    u = torch.nn.functional.normalize(torch.randn(64, device="cuda", generator=g), dim=0)
    qkv[:, :, 0] += 6.0 * u                      # every query has a strong component along u
    for j in range((S + 127) // 128):            # one spike key per 128-key tile, stronger in later tiles
        pos = min(j * 128 + 37, S - 1)
        qkv[:, pos, 1] = u * (8.0 * (j + 1))     # score ~ 6 * 8 (j+1) / 8 = 6 (j+1) after the 1/8 scale ... x log2e
    qkv = qkv.reshape(n_seq * S, 3 * H).to(tdt)
    out = torch.full((n_seq * S, H), float("nan"), device="cuda", dtype=tdt)
    ctx.call("vtq_attention_fwd", P(qkv), P(out), n_seq, S, heads, code, 0, ST())
    torch.cuda.synchronize()
    want = _attn_ref(qkv, n_seq, S, heads)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - want).abs().max().item()
    assert err < (3e-2 if dt == "bf16" else 4e-3), err
    # the growth between consecutive planted scores really exceeds the lazy-rescale threshold of 8 (exp2 domain)
    q = qkv.view(n_seq, S, 3, heads, 64)[0, 5, 0, 0].float()
    k0 = qkv.view(n_seq, S, 3, heads, 64)[0, 37, 1, 0].float()
    k1 = qkv.view(n_seq, S, 3, heads, 64)[0, 165, 1, 0].float()
    assert ((q @ k1) - (q @ k0)).item() * 0.125 * 1.4427 > 8.0


# ------------------------------------------------------------------------------------------ row-wise kernels
@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("rows", [1, 7, 1003])
def test_layernorm(ctx, dt, rows):
    code, tdt = DT[dt]
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, 768, device="cuda", generator=g) * 3 + 0.7
    w = torch.randn(768, device="cuda", generator=g)
    b = torch.randn(768, device="cuda", generator=g)
    out = torch.empty(rows, 768, device="cuda", dtype=tdt)
    ctx.call("vtq_layernorm", P(x), 0, P(w), P(b), 1e-6, rows, 768, P(out), code, ST())
    torch.cuda.synchronize()
    want = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-6)
    # identical up to the final 16-bit rounding
    assert (out.float() - want.to(tdt).float()).abs().max().item() <= (0.04 if dt == "bf16" else 0.005)
    assert (out.float() - want).abs().max().item() < (0.05 if dt == "bf16" else 0.006)


def test_layernorm_strided_rows_and_attention_row_limit(ctx):
    """The two ABI features the last-block pruning uses: strided LayerNorm rows, attention limited to leading rows."""
    g = torch.Generator(device="cuda").manual_seed(4)
    S, n_seq, H = 301, 3, 768
    x = torch.randn(n_seq * S, H, device="cuda", generator=g)
    w = torch.randn(H, device="cuda", generator=g); b = torch.randn(H, device="cuda", generator=g)
    out = torch.empty(n_seq, H, device="cuda", dtype=torch.float16)
    ctx.call("vtq_layernorm", P(x), S * H, P(w), P(b), 1e-6, n_seq, H, P(out), 0, ST())
    want = torch.nn.functional.layer_norm(x[::S], (H,), w, b, 1e-6)
    qkv = (torch.randn(n_seq * S, 3 * H, device="cuda", generator=g) * 1.5).half()
    att = torch.full((n_seq * S, H), 7.0, device="cuda", dtype=torch.float16)
    ctx.call("vtq_attention_fwd", P(qkv), P(att), n_seq, S, 12, 0, 1, ST())
    torch.cuda.synchronize()
    assert (out.float() - want).abs().max().item() < 0.006
    ref = _attn_ref(qkv, n_seq, S, 12).view(n_seq, S, H)
    got = att.view(n_seq, S, H)
    assert (got[:, :256].float() - ref[:, :256]).abs().max().item() < 4e-3    # first 256-row granule computed
    assert torch.all(got[:, 256:] == 7.0)                                     # rows beyond it untouched


def test_cast_rows(ctx):
    x = torch.randn(5, 768, device="cuda")
    for name, (code, tdt) in DT.items():
        out = torch.empty(5, 768, device="cuda", dtype=tdt)
        ctx.call("vtq_cast_rows", P(x), P(out), x.numel(), code, ST())
        torch.cuda.synchronize()
        assert torch.equal(out, x.to(tdt))


@pytest.mark.parametrize("n_extra,use_scales", [(0, False), (8, True)])
def test_embed_assemble_indices_bit_exact(ctx, n_extra, use_scales):
    n_seq, N, H = 3, 77, 768
    g = torch.Generator(device="cuda").manual_seed(3)
    proj = torch.randn(n_seq * N, H, device="cuda", generator=g)
    pos = torch.rand(n_seq * N, 2, device="cuda", generator=g)
    pos[0] = torch.tensor([0.0, 0.0]); pos[1] = torch.tensor([0.99999899, 0.99999899]); pos[2] = torch.tensor([1 / 24, 23 / 24])
    scales = torch.randint(0, 5, (n_seq * N,), device="cuda", generator=g).float() if use_scales else None
    pos_table = torch.randn(577, H, device="cuda", generator=g)
    scale_table = torch.randn(4, H, device="cuda", generator=g) if use_scales else None
    cls = torch.randn(H, device="cuda", generator=g)
    extra = torch.randn(n_extra, H, device="cuda", generator=g) if n_extra else None
    T = 1 + n_extra
    x = torch.full((n_seq, T + N, H), float("nan"), device="cuda")
    pidx = torch.zeros(n_seq * N, dtype=torch.int32, device="cuda")
    sidx = torch.zeros(n_seq * N, dtype=torch.int32, device="cuda")
    ctx.call("vtq_embed_assemble", P(proj), P(pos), P(scales), P(pos_table), 24, P(scale_table), 3 if use_scales else 0,
             P(cls), P(extra), n_extra, n_seq, N, H, P(x), P(pidx), P(sidx), ST())
    torch.cuda.synchronize()
    want_idx = patch_oracle.pos_index(pos.cpu().numpy(), 24)
    assert np.array_equal(pidx.cpu().numpy().astype(np.int64), want_idx)
    want = proj + pos_table[torch.from_numpy(want_idx).cuda()]
    if use_scales:
        want_s = patch_oracle.scale_index(scales.cpu().numpy(), 3)
        assert np.array_equal(sidx.cpu().numpy().astype(np.int64), want_s)
        want = want + scale_table[torch.from_numpy(want_s).cuda()]
    assert torch.equal(x[:, T:].reshape(-1, H), want)          # fp32 adds in the reference's order: bit-exact
    assert torch.equal(x[:, 0], (cls + pos_table[0]).expand(n_seq, H))
    if n_extra:
        assert torch.equal(x[:, 1:T], extra.expand(n_seq, n_extra, H))


def test_embed_requires_scales(ctx):
    from vtamiq_b200._lib import VtqError
    t = torch.zeros(8, 768, device="cuda")
    with pytest.raises(VtqError, match="scales is passed as None"):
        ctx.call("vtq_embed_assemble", P(t), P(t), None, P(t), 24, P(t), 3, P(t), None, 0, 1, 8, 768, P(t), None, None, ST())


def test_cls_diff(ctx):
    B, S, H = 5, 9, 768
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(2 * B, S, H, device="cuda", generator=g) * 2
    w = torch.randn(H, device="cuda", generator=g); b = torch.randn(H, device="cuda", generator=g)
    gamma = torch.rand(H, device="cuda", generator=g) + 0.5
    out = torch.empty(B, H, device="cuda")
    ctx.call("vtq_cls_diff", P(x), C.c_void_p(x.data_ptr() + B * S * H * 4), B, S, H, 0, P(w), P(b), 1e-6, P(gamma), P(out), ST())
    torch.cuda.synchronize()
    ln = torch.nn.functional.layer_norm(x[:, 0].cpu(), (H,), w.cpu(), b.cpu(), 1e-6)
    want = (ln[:B] - ln[B:]) * gamma.cpu()
    assert (out.cpu() - want).abs().max().item() < 5e-6


@pytest.mark.parametrize("B", [1, 5, 33])
def test_diffnet_head_matches_oracle(ctx, B):
    import vtamiq_b200
    torch.manual_seed(0)
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=1)).eval()
    synth.perturb_(m)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    eng = m.engine
    eng._ensure_ready()
    d = torch.randn(B, 768, generator=torch.Generator().manual_seed(B)) * 0.5
    want = vtamiq_oracle.diffnet_head(sd, vtamiq_oracle._cfg_from_state(sd), d)
    dd = d.cuda()
    q = torch.empty(B, device="cuda")
    ws = torch.empty(ctx.workspace_bytes(B, 768), dtype=torch.uint8, device="cuda")
    ctx.call("vtq_diffnet_head", P(dd), eng._tail_params, len(eng._tail_params), eng.num_rgs, eng.num_rcabs, 768,
             eng.ca_hidden, eng.head_hidden, B, P(q), P(ws), ST())
    torch.cuda.synchronize()
    assert (q.cpu() - want).abs().max().item() < 2e-5


# ------------------------------------------------------------------------------------------ patch gather
@pytest.mark.parametrize("case", ["single", "multi3", "odd2", "clamp"])
def test_patch_gather_bit_exact_vs_reference_golden(golden_dir, case):
    from vtamiq_b200 import extract_patches
    g = np.load(os.path.join(golden_dir, f"patches_{case}.npz"))
    tens = torch.stack([synth.to_tensor_normalized(g["ref_u8"]), synth.to_tensor_normalized(g["dist_u8"])]).cuda()
    smp = [g[f"samples_{i}"] for i in range(int(g["n_levels"]))]
    patches, pos, scales = extract_patches(tens, smp)
    torch.cuda.synchronize()
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), g["patches"].view(np.uint32))
    assert np.array_equal(pos.cpu().numpy().view(np.uint32), g["pos"].view(np.uint32))
    if "scales" in g.files:
        assert scales.dtype == torch.int32 and np.array_equal(scales.cpu().numpy(), g["scales"])
    else:
        assert scales is None


def test_patch_gather_full_size_vs_oracle():
    """cfg3-like: 1024x1024, 3 levels, 500 patches (380/96/24) — bit-exact against the numpy oracle."""
    from vtamiq_b200 import extract_patches
    rng = np.random.default_rng(7)
    tens = rng.standard_normal((2, 3, 1024, 1024)).astype(np.float32)
    smp = [synth.jittered_samples(rng, 1024 >> s, 1024 >> s, n) for s, n in enumerate((380, 96, 24))]
    smp[0][:, 0] = [0.0, 0.0]; smp[0][:, 1] = [1008.0, 1008.0]; smp[0][:, 2] = [1007.999999, 0.5]   # edges
    want_p, want_pos, want_s = patch_oracle.extract_patches(tens, smp)
    patches, pos, scales = extract_patches(torch.from_numpy(tens).cuda(), smp)
    torch.cuda.synchronize()
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), want_p.view(np.uint32))
    assert np.array_equal(pos.cpu().numpy().view(np.uint32), want_pos.view(np.uint32))
    assert np.array_equal(scales.cpu().numpy(), want_s)


def test_avgpool_bit_exact(ctx):
    x = torch.randn(5, 37, 50, device="cuda")
    out = torch.empty(5, 18, 25, device="cuda")
    ctx.call("vtq_avgpool2x2", P(x), P(out), 5, 37, 50, ST())
    y = torch.randn(4, 21, 33, device="cuda")       # odd width -> scalar-load variant
    out2 = torch.empty(4, 10, 16, device="cuda")
    ctx.call("vtq_avgpool2x2", P(y), P(out2), 4, 21, 33, ST())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), patch_oracle.avgpool2x2(x.cpu().numpy()).view(np.uint32))
    assert np.array_equal(out2.cpu().numpy().view(np.uint32), patch_oracle.avgpool2x2(y.cpu().numpy()).view(np.uint32))


@pytest.mark.parametrize("case", ["single", "multi3", "odd2"])
def test_patch_gather_from_uint8_bit_exact(golden_dir, case):
    """uint8 HWC source with the reference transform fused in (or applied on device in front of the pyramid)."""
    from vtamiq_b200 import extract_patches
    g = np.load(os.path.join(golden_dir, f"patches_{case}.npz"))
    u8 = torch.from_numpy(np.stack([g["ref_u8"], g["dist_u8"]])).cuda()      # (2, H, W, 3) uint8
    smp = [g[f"samples_{i}"] for i in range(int(g["n_levels"]))]
    patches, pos, scales = extract_patches(u8, smp)
    torch.cuda.synchronize()
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), g["patches"].view(np.uint32))
    assert np.array_equal(pos.cpu().numpy().view(np.uint32), g["pos"].view(np.uint32))
    if "scales" in g.files:
        assert np.array_equal(scales.cpu().numpy(), g["scales"])


# ------------------------------------------------------------------------------------------ K1, round 2 additions
def _oracle_from_u8(u8_pair, smp):
    """Reference arithmetic on decoded images: transform (to_tensor + normalize) then the numpy gather oracle."""
    tens = np.stack([synth.to_tensor_normalized(u8_pair[k]).numpy() for k in range(u8_pair.shape[0])])
    return patch_oracle.extract_patches(tens, smp)


@pytest.mark.parametrize("src", ["uint8", "fp32"])
def test_batched_gather_cfg2_shape_bit_exact_incl_16bit_operand(ctx, src):
    """The vector gather kernels (256-bit loads, register rotation, division-free uint8 transform) at the benchmark
    shape, one coordinate set per pair: fp32 patches / uv bit-exact against the oracle, and the 16-bit GEMM operand
    equal to the fp32 patches rounded once."""
    from vtamiq_b200 import extract_patches_batch
    B, H, W, N = 3, 384, 512, 500
    rng = np.random.default_rng(17)
    u8 = np.stack([np.stack([synth.make_pair(p, H, W, 0.1)[k] for p in range(B)]) for k in range(2)])   # (2,B,H,W,3)
    smp = np.stack([synth.jittered_samples(rng, H, W, N) for _ in range(B)])
    smp[0, :, 0] = [0.0, 0.0]; smp[0, :, 1] = [H - 16.0, W - 16.0]; smp[1, :, 2] = [367.99999, 0.25]     # edges
    if src == "uint8":
        images = torch.from_numpy(u8).cuda()
    else:
        images = torch.stack([torch.stack([synth.to_tensor_normalized(u8[k, p]) for p in range(B)]) for k in range(2)]).cuda()
    patches, pos, scales = extract_patches_batch(images, [torch.from_numpy(smp).cuda()])
    assert scales is None
    for p in range(B):
        want_p, want_pos, _ = _oracle_from_u8(u8[:, p], [smp[p]])
        assert np.array_equal(patches[:, p].cpu().numpy().view(np.uint32), want_p.view(np.uint32)), p
        assert np.array_equal(pos[:, p].cpu().numpy().view(np.uint32), want_pos.view(np.uint32)), p
    # 16-bit operand written by the same kernels (what Engine consumes)
    p16 = torch.empty(2 * B * N, 768, dtype=torch.float16, device="cuda")
    pos2 = torch.empty(2 * B * N, 2, device="cuda")
    s_dev = torch.from_numpy(smp).cuda()
    name = "vtq_patch_gather_u8" if src == "uint8" else "vtq_patch_gather"
    ctx.call(name, P(images), 2 * B, H, W, P(s_dev), B, N, 0, N, None, P(p16), 0, P(pos2), None, 0, ST())
    torch.cuda.synchronize()
    assert torch.equal(p16.view(2, B, N, 768), patches.reshape(2, B, N, 768).half())
    assert torch.equal(pos2.view(2, B, N, 2), pos)


def test_batched_multiscale_gather_from_uint8_fused_pyramid():
    """cfg3 shape from decoded images: level 0 gathered straight from uint8, level 1 = transform + 2x2 mean in one
    kernel (no fp32 level-0 image), level 2 pooled from level 1 — bit-exact against transform -> AvgPool2d chain."""
    from vtamiq_b200 import extract_patches_batch
    B, H, W = 2, 1024, 1024
    counts = (380, 96, 24)
    rng = np.random.default_rng(23)
    u8 = rng.integers(0, 256, size=(2, B, H, W, 3), dtype=np.uint8)
    smp = [np.stack([synth.jittered_samples(rng, H >> s, W >> s, n) for _ in range(B)]) for s, n in enumerate(counts)]
    patches, pos, scales = extract_patches_batch(torch.from_numpy(u8).cuda(), [torch.from_numpy(s).cuda() for s in smp])
    for p in range(B):
        want_p, want_pos, want_s = _oracle_from_u8(u8[:, p], [s[p] for s in smp])
        assert np.array_equal(patches[:, p].cpu().numpy().view(np.uint32), want_p.view(np.uint32)), p
        assert np.array_equal(pos[:, p].cpu().numpy().view(np.uint32), want_pos.view(np.uint32)), p
        assert np.array_equal(scales[:, p].cpu().numpy().astype(np.int64), want_s.astype(np.int64)), p


def test_u8_transform_and_pool_kernels_bit_exact(ctx):
    """vtq_normalize_u8 (vector and scalar shapes), vtq_avgpool2x2 (vector) and vtq_avgpool2x2_u8 against torch."""
    g = torch.Generator().manual_seed(3)
    for (n, H, W) in [(3, 40, 64), (2, 31, 33)]:
        u8 = torch.randint(0, 256, (n, H, W, 3), generator=g, dtype=torch.uint8)
        want = torch.stack([synth.to_tensor_normalized(u8[i].numpy()) for i in range(n)])
        out = torch.empty(n, 3, H, W, device="cuda")
        ctx.call("vtq_normalize_u8", P(u8.cuda()), P(out), n, H, W, ST())
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), want), (n, H, W)
    x = torch.randn(6, 64, 96, generator=g)
    out = torch.empty(6, 32, 48, device="cuda")
    ctx.call("vtq_avgpool2x2", P(x.cuda()), P(out), 6, 64, 96, ST())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), patch_oracle.avgpool2x2(x.numpy()).view(np.uint32))
    u8 = torch.randint(0, 256, (3, 50, 72, 3), generator=g, dtype=torch.uint8)
    want = torch.stack([synth.to_tensor_normalized(u8[i].numpy()) for i in range(3)]).numpy()
    out = torch.empty(3, 3, 25, 36, device="cuda")
    ctx.call("vtq_avgpool2x2_u8", P(u8.cuda()), P(out), 3, 50, 72, ST())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), patch_oracle.avgpool2x2(want).view(np.uint32))


def test_out_of_range_coordinates_raise_index_error(ctx):
    """torch's advanced indexing raises IndexError in the reference gather (patch_sampling.py:531-545); here the
    kernels clamp (no out-of-bounds read) and raise a flag that the host mirror turns into IndexError."""
    from vtamiq_b200 import check_coordinates, extract_patches
    H, W = 64, 96
    tens = torch.randn(2, 3, H, W, device="cuda")
    good = np.array([[0.0, 10.5, 48.0], [0.0, 20.25, 80.0]])
    check_coordinates(ctx, sync=True)                      # clean slate
    extract_patches(tens, [good])
    for bad in ([[49.0], [0.0]], [[0.0], [81.0]], [[-1.0], [5.0]], [[float("nan")], [5.0]]):
        with pytest.raises(IndexError, match="outside the image"):
            extract_patches(tens, [np.array(bad, dtype=np.float64)])
    extract_patches(tens, [good])                          # the flag does not stick


def test_tensor_map_cache_is_hit_in_steady_state(ctx):
    """TMA descriptors are encoded once per distinct (pointer, shape, box) and re-used on later launches."""
    M, N, K = 512, 768, 768
    A = torch.randn(M, K, device="cuda").half(); W = torch.randn(N, K, device="cuda").half()
    b = torch.zeros(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ctx.call("vtq_gemm", P(A), 0, P(W), P(b), M, N, K, 0, 0, P(out), 0, None, ST())
    h0, m0 = ctx.tensor_map_stats()
    for _ in range(3):
        ctx.call("vtq_gemm", P(A), 0, P(W), P(b), M, N, K, 0, 0, P(out), 0, None, ST())
    torch.cuda.synchronize()
    h1, m1 = ctx.tensor_map_stats()
    assert m1 == m0 and h1 - h0 == 9      # three descriptors per launch, all from the cache


# ---- the non-default branches of get_iqa_patches (patch_sampling.py:506-508, :530-531/:561, :46-222/:603-605) ----
class _ReplaySampler:
    """Stands in for the reference's PatchSampler on a box without the reference: returns the draws the real sampler
    made when the fixture was generated (tests/golden/make_golden.py::patches_branch_case) and records the weight map
    it is handed at every draw."""

    def __init__(self, draws, diff0=None):
        self.draws, self.diff0, self.seen, self.i = list(draws), diff0, [], 0

    def compute_diff(self, imgs):
        return None if self.diff0 is None else np.array(self.diff0, copy=True)

    def get_sample_params(self, h, w, ho, wo, diff=None, num_samples=1, debug=False):
        self.seen.append(None if diff is None else np.array(diff, copy=True))
        d = self.draws[self.i]
        self.i += 1
        assert d.shape == (2, num_samples), (d.shape, num_samples)
        return d


def _branch_fixture(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"patches_{name}.npz"))
    tens = [synth.to_tensor_normalized(g["ref_u8"]).cuda(), synth.to_tensor_normalized(g["dist_u8"]).cuda()]
    draws = [g[f"draw_{i}"] for i in range(int(g["n_draws"]))]
    return g, tens, draws


def test_get_iqa_patches_unaligned_matches_reference(golden_dir):
    """use_aligned_patches=False: ref and dist get their own coordinate sets (two draws per level)."""
    from vtamiq_b200 import get_iqa_patches
    g, tens, draws = _branch_fixture(golden_dir, "unaligned")
    smp = _ReplaySampler(draws)
    patches, pos, scales = get_iqa_patches((g["ref_u8"], g["dist_u8"]), tens, int(g["N"]), 16, smp,
                                           int(g["n_scales_requested"]), scale_num_samples_ratio=float(g["ratio"]),
                                           use_aligned_patches=False, random_seed=int(g["seed"]))
    torch.cuda.synchronize()
    assert smp.i == len(draws) == 4
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), g["patches"].view(np.uint32))
    assert np.array_equal(pos.cpu().numpy().view(np.uint32), g["pos"].view(np.uint32))
    assert np.array_equal(scales.cpu().numpy(), g["scales"])
    assert not np.array_equal(g["pos"][0], g["pos"][1])      # the two images really use different positions


def test_get_iqa_patches_difference_weighted_matches_reference(golden_dir):
    """Difference-weighted sampler: compute_diff's map reaches the sampler at level 0 and, 2x mean-pooled, at level 1
    (bit-identical to the reference's pooling); extraction bit-exact."""
    from vtamiq_b200 import get_iqa_patches
    g, tens, draws = _branch_fixture(golden_dir, "weighted")
    smp = _ReplaySampler(draws, diff0=g["diff0"])
    patches, pos, scales = get_iqa_patches((g["ref_u8"], g["dist_u8"]), tens, int(g["N"]), 16, smp,
                                           int(g["n_scales_requested"]), scale_num_samples_ratio=float(g["ratio"]),
                                           random_seed=int(g["seed"]))
    torch.cuda.synchronize()
    assert len(smp.seen) == 2
    for i, w in enumerate(smp.seen):
        assert w.dtype == g[f"weight_{i}"].dtype and np.array_equal(w, g[f"weight_{i}"])
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), g["patches"].view(np.uint32))
    assert np.array_equal(pos.cpu().numpy().view(np.uint32), g["pos"].view(np.uint32))
    assert np.array_equal(scales.cpu().numpy(), g["scales"])


def test_get_iqa_patches_random_slot_order(golden_dir):
    """randomize_patch_scale_order=True: slot perm[i] receives the i-th patch of the scale-ordered sequence, perm =
    the first thing drawn from numpy's RNG after seeding (patch_sampling.py:506-508, :587-590).  The reference itself
    raises on this branch under torch 2.11 (index_put of float64 into float32, :534), so the expectation is the
    scale-ordered result of the same draws, permuted."""
    from vtamiq_b200 import get_iqa_patches
    g = np.load(os.path.join(golden_dir, "patches_multi3.npz"))
    tens = [synth.to_tensor_normalized(g["ref_u8"]).cuda(), synth.to_tensor_normalized(g["dist_u8"]).cuda()]
    draws = [g[f"samples_{i}"] for i in range(int(g["n_levels"]))]
    N, seed = int(g["N"]), 31
    args = ((g["ref_u8"], g["dist_u8"]), tens, N, 16)
    kw = dict(scale_num_samples_ratio=float(g["ratio"]), random_seed=seed)
    plain = get_iqa_patches(*args, _ReplaySampler(draws), int(g["n_scales_requested"]), **kw)
    mixed = get_iqa_patches(*args, _ReplaySampler(draws), int(g["n_scales_requested"]),
                            randomize_patch_scale_order=True, **kw)
    torch.cuda.synchronize()
    perm = np.random.RandomState(seed).permutation(N)
    for a, b in zip(plain, mixed):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        assert np.array_equal(b[:, perm], a)
    assert not np.array_equal(mixed[2].cpu().numpy(), plain[2].cpu().numpy())   # scale ids are no longer sorted
    assert np.array_equal(plain[0].cpu().numpy().view(np.uint32), g["patches"].view(np.uint32))


def test_device_sampler_kernel_has_the_reference_law(golden_dir):
    """vtq_sample_grid (one launch per level, Philox-keyed permutation of the grid cells + jitter) against raw draws of
    the reference's default sampler (tests/golden/sampler_draws.npz): same support, one sample per distinct grid cell,
    same jitter range, marginals a two-sample KS test cannot separate; reproducible from the generator state."""
    import scipy.stats
    from vtamiq_b200.patch_sampling import perturbed_grid_samples, sample_batch
    z = np.load(os.path.join(golden_dir, "sampler_draws.npz"))
    g = torch.Generator(device="cuda").manual_seed(3)
    for name in ("cfg2", "small", "tall"):
        ref = z[name].astype(np.float64)                       # (draws, 2, n)
        h, w, n = (int(v) for v in z[name + "_hwn"])
        ours = perturbed_grid_samples(ref.shape[0], h, w, 16, 16, n, device="cuda", generator=g).cpu().numpy()
        assert ours.shape == ref.shape and ours.dtype == np.float64
        width = max(int(np.ceil(np.sqrt(n / (h / w)))), 1)
        height = int(np.ceil(width * h / w))
        for smp in (ours, ref):
            assert smp[:, 0].min() >= 0 and smp[:, 0].max() <= h - 16 and smp[:, 1].min() >= 0 and smp[:, 1].max() <= w - 16
            cy = np.minimum(np.floor(smp[:, 0] / (h - 16) * height), height - 1)
            cx = np.minimum(np.floor(smp[:, 1] / (w - 16) * width), width - 1)
            cell = (cy * width + cx).astype(int)
            assert all(len(set(row)) == n for row in cell)     # n distinct grid cells per image
            off_y = smp[:, 0] / (h - 16) * height - cy - 0.5   # jitter inside the cell, |.| <= 2 * 0.2
            off_x = smp[:, 1] / (w - 16) * width - cx - 0.5
            assert np.abs(off_y).max() <= 0.4 + 1e-6 and np.abs(off_x).max() <= 0.4 + 1e-6
        for axis in (0, 1):
            p = scipy.stats.ks_2samp(ours[:, axis].ravel(), ref[:, axis].ravel()).pvalue
            assert p > 1e-3, (name, axis, p)
        # images of one batch get different draws; the same generator state reproduces the batch
        assert not np.array_equal(ours[0], ours[1])
    a = perturbed_grid_samples(4, 384, 512, 16, 16, 500, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    b = perturbed_grid_samples(4, 384, 512, 16, 16, 500, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    assert torch.equal(a, b)
    levels = sample_batch(3, 1024, 1024, 500, 16, 3, 2.0, device="cuda", generator=g)
    assert [tuple(t.shape) for t in levels] == [(3, 2, 380), (3, 2, 96), (3, 2, 24)]
    assert float(levels[2][:, 0].max()) <= 256 - 16
    big = perturbed_grid_samples(2, 2160, 3840, 16, 16, 5000, device="cuda", generator=g)      # cfg4: 8192-key sort
    assert big.shape == (2, 2, 5000) and float(big[:, 0].max()) <= 2160 - 16 and float(big[:, 1].max()) <= 3840 - 16
