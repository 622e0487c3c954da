"""Deterministic synthetic inputs shared by the golden generator, the tests, smoke() and bench.py.

Image content follows SURVEY.md §8d: per pair p, ``rng = default_rng(1000+p)``; the reference image is either
uniform noise or 1/f-filtered noise; the distorted image is the reference plus a graded, tie-free distortion
(additive Gaussian noise on even p, Gaussian blur on odd p), re-quantised to uint8.
"""
from __future__ import annotations

import numpy as np
import torch


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """Separable Gaussian blur, reflect padding, float64 (no scipy: keeps fixtures machine-independent)."""
    r = max(1, int(np.ceil(3 * sigma)))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    out = img.astype(np.float64)
    for axis in (0, 1):
        pad = [(0, 0)] * out.ndim
        pad[axis] = (r, r)
        p = np.pad(out, pad, mode="reflect")
        acc = np.zeros_like(out)
        for i, kv in enumerate(k):
            sl = [slice(None)] * out.ndim
            sl[axis] = slice(i, i + out.shape[axis])
            acc += kv * p[tuple(sl)]
        out = acc
    return out


def make_pair(p: int, H: int, W: int, level: float, family: str = "auto"):
    """uint8 (H,W,3) reference and distorted images for pair index p; ``level`` in (0,1] grades the distortion."""
    rng = np.random.default_rng(1000 + p)
    if family == "auto":
        family = "pink" if (p // 2) % 2 else "white"
    if family == "white":
        ref = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    else:
        spec = np.fft.rfft2(rng.standard_normal((H, W, 3)), axes=(0, 1))
        fy = np.fft.fftfreq(H)[:, None]
        fx = np.fft.rfftfreq(W)[None, :]
        f = np.sqrt(fy * fy + fx * fx)
        f[0, 0] = 1.0
        img = np.fft.irfft2(spec / f[..., None], s=(H, W), axes=(0, 1))
        img = (img - img.min()) / (img.max() - img.min())
        ref = np.round(img * 255).astype(np.uint8)
    reff = ref.astype(np.float64) / 255.0
    if p % 2 == 0:
        dist = reff + rng.standard_normal(reff.shape) * level
    else:
        dist = _blur(reff, 20.0 * level)
    dist = np.clip(np.round(dist * 255), 0, 255).astype(np.uint8)
    return ref, dist


def graded_levels(B: int, seed: int = 0) -> np.ndarray:
    """Tie-free distortion levels, geomspace(4/255, 96/255, B) in a fixed permutation."""
    lv = np.geomspace(4 / 255, 96 / 255, B)
    return lv[np.random.default_rng(seed).permutation(B)]


def to_tensor_normalized(u8: np.ndarray) -> torch.Tensor:
    """uint8 (H,W,3) -> fp32 (3,H,W) in [-1,1]: to_tensor then normalize(mean .5, std .5), i.e. the reference's
    transform_img (data/utils.py:76,:94; patch_datasets.py:51-52) — same fp32 op order (div 255, sub .5, div .5)."""
    t = torch.from_numpy(np.ascontiguousarray(u8)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    return t.sub_(0.5).div_(0.5)


def jittered_samples(rng: np.random.Generator, h: int, w: int, n: int, patch: int = 16) -> np.ndarray:
    """float64 (2,n) top-left coordinates in [0,h-patch] x [0,w-patch] on a jittered grid.  Input generator for
    tests/bench on machines without the reference sampler; NOT a restatement of stratified_grid_sampling."""
    aspect = h / w
    cols = max(1, int(np.ceil(np.sqrt(n / aspect))))
    rows = max(1, int(np.ceil(n / cols)))
    idx = rng.permutation(rows * cols)[:n]
    gy, gx = idx // cols, idx % cols
    y = (gy + rng.random(n)) / rows * (h - patch)
    x = (gx + rng.random(n)) / cols * (w - patch)
    return np.stack([np.clip(y, 0, h - patch), np.clip(x, 0, w - patch)]).astype(np.float64)


def perturb_(model: torch.nn.Module, seed: int = 1) -> None:
    """Make every parameter class matter: random-init leaves all biases 0, LayerNorm at (1,0) and every LayerScale
    at 1, which would hide bias/gamma/beta bugs.  Deterministic given the (identical) parameter order."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("gamma"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias") and name.startswith("transformer."):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))


def state_hash(sd) -> str:
    import hashlib
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def select_separated(scores: np.ndarray, k: int, min_gap: float) -> np.ndarray:
    """Indices of k pairs whose (reference) scores are pairwise >= min_gap apart, spread over the score range.
    SRCC >= 0.9999 at batch 32 tolerates zero rank swaps, so the measurement batch must be tie-free and
    well separated; selection uses the reference scores only."""
    order = np.argsort(scores)
    chosen = [order[0]]
    for i in order[1:]:
        if scores[i] - scores[chosen[-1]] >= min_gap:
            chosen.append(i)
    if len(chosen) < k:
        raise ValueError(f"only {len(chosen)} of the requested {k} pairs are {min_gap} apart")
    pick = np.round(np.linspace(0, len(chosen) - 1, k)).astype(int)
    return np.sort(np.array(chosen)[pick])


def synthetic_vit_npz(seed: int = 0, pos_tokens: int = 577, hidden: int = 768, mlp: int = 3072, heads: int = 12,
                      layers: int = 12, patch: int = 16) -> dict:
    """A JAX-format ViT checkpoint (the key/shape layout of Google's ViT-B_16.npz that
    ``VisionTransformer.load_from`` reads, transformer.py:287-325,:428-455,:643-668) filled with seeded noise.
    ``pos_tokens`` = 197 gives a 14x14 positional grid, which makes load_from take its ndimage.zoom resize branch."""
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s, dtype=np.float32)
    d = hidden // heads
    w = {"embedding/kernel": f(patch, patch, 3, hidden), "embedding/bias": f(hidden), "cls": f(1, 1, hidden),
         "Transformer/posembed_input/pos_embedding": f(1, pos_tokens, hidden),
         "Transformer/encoder_norm/scale": f(hidden), "Transformer/encoder_norm/bias": f(hidden)}
    for i in range(layers):
        r = f"Transformer/encoderblock_{i}/"
        for n in ("query", "key", "value"):
            w[r + f"MultiHeadDotProductAttention_1/{n}/kernel"] = f(hidden, heads, d)
            w[r + f"MultiHeadDotProductAttention_1/{n}/bias"] = f(heads, d)
        w[r + "MultiHeadDotProductAttention_1/out/kernel"] = f(heads, d, hidden)
        w[r + "MultiHeadDotProductAttention_1/out/bias"] = f(hidden)
        w[r + "MlpBlock_3/Dense_0/kernel"] = f(hidden, mlp)
        w[r + "MlpBlock_3/Dense_0/bias"] = f(mlp)
        w[r + "MlpBlock_3/Dense_1/kernel"] = f(mlp, hidden)
        w[r + "MlpBlock_3/Dense_1/bias"] = f(hidden)
        for n in ("LayerNorm_0", "LayerNorm_2"):
            w[r + n + "/scale"] = f(hidden)
            w[r + n + "/bias"] = f(hidden)
    return w


def grad_probe(t) -> np.ndarray:
    """Compact fingerprint of a gradient tensor for fixtures: [sum, sum|.|, 24 strided samples]."""
    a = t.detach().cpu().double().reshape(-1).numpy()
    idx = np.unique(np.linspace(0, a.size - 1, 24).round().astype(np.int64))
    pad = np.full(24 - idx.size, np.nan)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx], pad])
