"""Device-side multi-scale patch extraction — host mirror of the gather half of
``get_iqa_patches`` (reference data/patch_sampling.py:450-613).

What stays on the host, unchanged: WHERE to sample (``PatchSampler.get_sample_params`` →
``stratified_grid_sampling``, patch_sampling.py:46-395, numpy RNG).  What moves to the GPU: the gather
closure (:529-545), the chained 2x ``AvgPool2d`` pyramid (:552,:600), uv normalisation (:559-568) and the
scale ids (:572-574), through ``vtq_patch_gather`` / ``vtq_avgpool2x2``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import VTQ_F16, get_context
from .engine import _ptr, _stream

DEFAULT_NUM_SAMPLES_RATIO = 1.7


def compute_patch_num_scales(patch_num_scales, h, w, ho, wo):
    """How many pyramid levels the image supports (patch_sampling.py:398-411)."""
    if patch_num_scales <= 1:
        return 1
    side = max(ho, wo)
    dim, possible = min(h, w), 0
    while dim > 1:
        possible += 1
        dim = (dim - side) / 2
    return max(1, min(possible - 1, patch_num_scales))


def compute_num_patches_per_scale(patch_count, patch_num_scales, scale_num_samples_ratio):
    """Patch budget per level, coarsest level first (patch_sampling.py:427-447): geometric weights
    2^(ratio*i), rounded up, then trimmed from the fine end so the total is exactly ``patch_count``."""
    counts = 2 ** (scale_num_samples_ratio * np.arange(patch_num_scales))
    counts = np.ceil(counts * patch_count / np.sum(counts)).astype(int)
    running = np.cumsum(counts)
    for i in range(patch_num_scales):
        if patch_count <= running[i]:
            counts[i] -= running[i] - patch_count
            counts[i + 1:] = 0
            break
    return counts


def perturbed_grid_samples(batch: int, h: int, w: int, ho: int, wo: int, num_samples: int, *, device="cuda",
                           generator: torch.Generator | None = None, perturbed_amount: float = 0.2) -> torch.Tensor:
    """Top-left patch coordinates for ``batch`` images at once, drawn ON THE DEVICE with the law of the reference's
    default sampler: ``PatchSampler(grid_type=GRID_TYPE_PERTURBED_SIMPLE)`` →
    ``stratified_grid_sampling`` (patch_sampling.py:236-237, :308-327, :362-376).  That mode is one cell covering the
    image: a regular ``height x width`` grid with ``width = ceil(sqrt(n / (h/w)))``, ``height = ceil(width * h/w)``,
    ``n`` DISTINCT grid points chosen uniformly, each jittered by U(-2a, 2a) cells (a = perturbed_amount), moved to
    the cell centre, clipped to [0, 1] and scaled to [0, h-ho] x [0, w-wo].

    Not bit-reproducible against numpy's RNG stream — the parity definition is statistical (same support, same
    marginals; tests/test_host_logic.py compares against draws of the reference itself).  Returns float64
    (batch, 2, num_samples), row 0 = y, the layout ``forward_from_images`` takes.
    """
    if num_samples < 1:
        raise ValueError("num_samples must be positive")
    dev = torch.device(device)
    if dev.type == "cuda":
        # one kernel launch per level (vtq_sample_grid: Philox-keyed permutation of the grid cells sorted in shared
        # memory + jitter); the Philox key is drawn with the caller's generator and never leaves the device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        dev = torch.device("cuda", idx)
        key = torch.randint(-2 ** 62, 2 ** 62, (2,), dtype=torch.int64, device=dev, generator=generator)
        out = torch.empty(batch, 2, num_samples, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            get_context(idx).call("vtq_sample_grid", _ptr(key), batch, h, w, ho, wo, num_samples,
                                  float(perturbed_amount), _ptr(out), _stream(dev))
        return out
    # host tensors (CPU-side tests of the law): the same law as tensor expressions
    aspect = h / w
    width = max(int(np.ceil(np.sqrt(num_samples / aspect))), 1)
    height = int(np.ceil(width * aspect))
    cells = height * width                       # >= num_samples by construction
    kw = dict(device=device, generator=generator)
    # n distinct grid points per image: the first n entries of a uniform random permutation
    picks = torch.rand(batch, cells, **kw).argsort(dim=1)[:, :num_samples]
    gy = (picks % height).double()               # the reference flattens its (2, width, height) grid this way
    gx = (picks // height).double()
    jitter = (2.0 * torch.rand(batch, 2, num_samples, dtype=torch.float64, **kw) - 1.0) * (2.0 * perturbed_amount)
    py = ((gy + jitter[:, 0]) / height + 0.5 / height).clamp_(0.0, 1.0)
    px = ((gx + jitter[:, 1]) / width + 0.5 / width).clamp_(0.0, 1.0)
    return torch.stack([py * (h - ho), px * (w - wo)], dim=1)


def sample_batch(batch: int, h: int, w: int, patch_count: int, patch_dim: int = 16, patch_num_scales: int = 1,
                 scale_num_samples_ratio: float = DEFAULT_NUM_SAMPLES_RATIO, *, device="cuda",
                 generator: torch.Generator | None = None):
    """Per-scale coordinate sets for a whole batch, ready for ``VTAMIQ.forward_from_images(images, samples)``:
    the level count and the per-level patch budget follow the reference (patch_sampling.py:398-411, :427-447; finest
    level first, level s drawn in the (h >> s, w >> s) image like :575-600)."""
    n_scales = compute_patch_num_scales(patch_num_scales, h, w, patch_dim, patch_dim)
    counts = compute_num_patches_per_scale(patch_count, n_scales, scale_num_samples_ratio)[::-1]
    out, total = [], 0
    for lvl, n in enumerate(counts):
        out.append(perturbed_grid_samples(batch, h >> lvl, w >> lvl, patch_dim, patch_dim, int(n), device=device,
                                          generator=generator))
        total += int(n)
        if total >= patch_count:   # the reference stops as soon as the budget is spent (:606-607)
            break
    return out


def _as_call(ctx_or_call):
    return ctx_or_call.call if hasattr(ctx_or_call, "call") else ctx_or_call


def _pyramid(ctx, level0: torch.Tensor, num_levels: int, levels=None):
    """[planes..., H, W] fp32 -> list of levels; level s+1 = 2x2 mean of level s (floor mode).  ``levels`` may hold
    already-built leading levels (level 0 may be None when it was never materialised)."""
    levels = [level0] if levels is None else list(levels)
    dev = levels[-1].device
    while len(levels) < num_levels:
        src = levels[-1]
        H, W = src.shape[-2:]
        dst = torch.empty(*src.shape[:-2], H // 2, W // 2, dtype=torch.float32, device=dev)
        _as_call(ctx)("vtq_avgpool2x2", _ptr(src), _ptr(dst), src.numel() // (H * W), H, W, _stream(dev))
        levels.append(dst)
    return levels


def _normalize_u8(ctx, images_u8: torch.Tensor) -> torch.Tensor:
    """uint8 (..., H, W, 3) -> normalised fp32 (..., 3, H, W) with the reference's transform arithmetic."""
    lead, (H, W) = images_u8.shape[:-3], images_u8.shape[-3:-1]
    out = torch.empty(*lead, 3, H, W, dtype=torch.float32, device=images_u8.device)
    _as_call(ctx)("vtq_normalize_u8", _ptr(images_u8), _ptr(out), images_u8.numel() // (3 * H * W), H, W,
                  _stream(images_u8.device))
    return out


def _u8_levels(ctx, images_u8: torch.Tensor, num_levels: int):
    """Pyramid of decoded uint8 (..., H, W, 3) images WITHOUT an fp32 level 0: level 0 stays uint8 (gathered with the
    transform fused), level 1 = transform + 2x2 mean in one pass, coarser levels from level 1.
    Returns [None | fp32 level0, level1, ...] (level 0 is only materialised when W % 8 != 0)."""
    lead, (H, W) = images_u8.shape[:-3], images_u8.shape[-3:-1]
    if num_levels == 1:
        return [None]
    if W % 8 != 0:
        f32 = _normalize_u8(ctx, images_u8)
        return _pyramid(ctx, f32, num_levels)
    l1 = torch.empty(*lead, 3, H // 2, W // 2, dtype=torch.float32, device=images_u8.device)
    _as_call(ctx)("vtq_avgpool2x2_u8", _ptr(images_u8), _ptr(l1), images_u8.numel() // (3 * H * W), H, W,
                  _stream(images_u8.device))
    return _pyramid(ctx, None, num_levels, levels=[None, l1])


def check_coordinates(ctx, sync: bool):
    """Raise IndexError (as torch's advanced indexing does in the reference gather, patch_sampling.py:531-545) when a
    gather launch saw a patch origin outside [0, H-16] x [0, W-16].  The kernels clamp such origins (no out-of-bounds
    read) and raise a flag in pinned host memory; ``sync`` waits for the launches queued so far, otherwise the flag
    reflects the launches that have already finished (checked again at the next call)."""
    if sync:
        torch.cuda.current_stream(torch.device("cuda", ctx.device)).synchronize()
    if ctx.coord_status(reset=True):
        raise IndexError("vtamiq_b200 patch gather: a sampled patch origin lies outside the image "
                         "(valid range [0, H-16] x [0, W-16]); the reference's gather raises IndexError here")


def _check_samples(smp, B, n=None):
    if smp.dtype != torch.float64 or smp.dim() != 3 or smp.shape[0] != B or smp.shape[1] != 2 or \
            (n is not None and smp.shape[2] != n):
        raise ValueError("samples[s] must be float64 (B, 2, n_s)")


def gather_into_workspace(eng, ws, images: torch.Tensor, samples, validate: str = "lazy"):
    """Fill ws.patches16 / ws.pos / ws.scales for ``Engine.run`` from images + per-scale coordinates.
    images: (2,B,3,H,W) fp32 normalised, or (2,B,H,W,3) uint8 as decoded (transform fused into the gather).
    validate: "lazy" (default: out-of-range coordinates of EARLIER, finished launches raise IndexError now; this
    call's are reported by the next call or by ``check_coordinates``), "sync" (wait and raise for this call),
    "off"."""
    dev = eng.device
    if images.device != dev:
        raise ValueError(f"images must live on the model's device ({dev}), got {images.device}")
    samples = [s if s.device == dev else s.to(dev) for s in samples]   # a host pointer must never reach a launch
    if validate != "off":
        check_coordinates(eng.ctx, sync=False)
    st = _stream(dev)
    use_scales = eng.scale_table is not None
    sc_ptr = _ptr(ws.scales) if use_scales else None
    pyr = lambda name, *a: eng._call("pyramid", name, *a)       # tagged launches (bench.py's per-kernel table)
    gat = lambda name, *a: eng._call("patch_gather", name, *a)
    u8 = images.dtype == torch.uint8
    if u8:
        if images.dim() != 5 or images.shape[0] != 2 or images.shape[4] != 3:
            raise ValueError("uint8 images must be (2, B, H, W, 3)")
        images = images.contiguous()
        B = images.shape[1]
        levels = _u8_levels(pyr, images, len(samples))
    else:
        if images.dim() != 5 or images.shape[0] != 2 or images.shape[2] != 3:
            raise ValueError("images must be (2, B, 3, H, W)")
        if images.dtype != torch.float32 or not images.is_contiguous():
            images = images.to(torch.float32).contiguous()
        B = images.shape[1]
        levels = _pyramid(pyr, images, len(samples))
    off = 0
    for s, (lvl, smp) in enumerate(zip(levels, samples)):
        _check_samples(smp, B)
        smp = smp.contiguous()
        n = smp.shape[2]
        if lvl is None:      # level 0 straight from the decoded image
            H, W = images.shape[2], images.shape[3]
            gat("vtq_patch_gather_u8", _ptr(images), 2 * B, H, W, _ptr(smp), B, n, off, ws.N, None,
                _ptr(ws.patches16), eng.vtq16, _ptr(ws.pos), sc_ptr, s, st)
        else:
            H, W = lvl.shape[-2:]
            gat("vtq_patch_gather", _ptr(lvl), 2 * B, H, W, _ptr(smp), B, n, off, ws.N, None,
                _ptr(ws.patches16), eng.vtq16, _ptr(ws.pos), sc_ptr, s, st)
        off += n
    if off != ws.N:
        raise ValueError("sample counts do not add up to the workspace's patch count")
    if validate == "sync":
        check_coordinates(eng.ctx, sync=True)


def extract_patches_batch(images: torch.Tensor, samples, with_scales: bool | None = None, validate: str = "sync"):
    """Reference-format model inputs for a whole batch of pairs in one pass (what a DataLoader over
    ``get_iqa_patches`` would collate, train.py:252-267).

    images  : (2, B, 3, H, W) fp32 normalised or (2, B, H, W, 3) uint8 as decoded, on a CUDA device
    samples : list over scales of float64 (B, 2, n_s) top-left coordinates (ref and dist share them)
    returns (patches (2,B,N,3,16,16) fp32, pos (2,B,N,2) fp32, scales (2,B,N) fp32 | None): index [0] is the ref
    block, [1] the dist block — ``model((patches[0], patches[1]), (pos[0], pos[1]), (scales[0], scales[1]))``.
    """
    if images.device.type != "cuda":
        raise RuntimeError("extract_patches_batch runs on the GPU only (no CPU path)")
    dev = images.device
    ctx = get_context(dev.index if dev.index is not None else torch.cuda.current_device())
    u8 = images.dtype == torch.uint8
    if u8:
        if images.dim() != 5 or images.shape[0] != 2 or images.shape[4] != 3:
            raise ValueError("uint8 images must be (2, B, H, W, 3)")
        images = images.contiguous()
    else:
        if images.dim() != 5 or images.shape[0] != 2 or images.shape[2] != 3:
            raise ValueError("images must be (2, B, 3, H, W)")
        if images.dtype != torch.float32 or not images.is_contiguous():
            images = images.to(torch.float32).contiguous()
    B = images.shape[1]
    samples = [s.to(dev).contiguous() for s in samples]
    N = int(sum(s.shape[-1] for s in samples))
    use_scales = (len(samples) > 1) if with_scales is None else with_scales
    patches = torch.empty(2, B, N, 3, 16, 16, dtype=torch.float32, device=dev)
    pos = torch.empty(2, B, N, 2, dtype=torch.float32, device=dev)
    scales = torch.empty(2, B, N, dtype=torch.float32, device=dev) if use_scales else None
    with torch.cuda.device(dev):
        if validate != "off":
            check_coordinates(ctx, sync=False)
        levels = _u8_levels(ctx, images, len(samples)) if u8 else _pyramid(ctx, images, len(samples))
        off = 0
        for s, (lvl, smp) in enumerate(zip(levels, samples)):
            _check_samples(smp, B)
            n = smp.shape[2]
            if lvl is None:
                H, W = images.shape[2], images.shape[3]
                ctx.call("vtq_patch_gather_u8", _ptr(images), 2 * B, H, W, _ptr(smp), B, n, off, N, _ptr(patches), None,
                         VTQ_F16, _ptr(pos), _ptr(scales), s, _stream(dev))
            else:
                H, W = lvl.shape[-2:]
                ctx.call("vtq_patch_gather", _ptr(lvl), 2 * B, H, W, _ptr(smp), B, n, off, N, _ptr(patches), None,
                         VTQ_F16, _ptr(pos), _ptr(scales), s, _stream(dev))
            off += n
        if validate == "sync":
            check_coordinates(ctx, sync=True)
    return patches, pos, scales


def extract_patches(tensors: torch.Tensor, samples, patch_dim: int = 16, with_scales: bool | None = None):
    """Reference-format outputs for K images sharing one coordinate set (aligned patches).

    tensors : (K, 3, H, W) fp32 CUDA tensor (K=2 for FR-IQA: ref, dist)
    samples : list over scales of float64 arrays/tensors (2, n_s), or (K, 2, n_s) = one coordinate set per image
    returns (patches (K,N,3,P,P) fp32, pos (K,N,2) fp32, scales (K,N) int32 | None) — bit-identical to
    ``get_iqa_patches(...)`` run with the same coordinates.
    """
    if patch_dim != 16:
        raise ValueError("the gather kernel is specialised for 16x16 patches")
    if tensors.device.type != "cuda":
        raise RuntimeError("extract_patches runs on the GPU only (no CPU path)")
    ctx = get_context(tensors.device.index if tensors.device.index is not None else torch.cuda.current_device())
    u8 = tensors.dtype == torch.uint8     # (K, H, W, 3) as decoded: the reference transform is applied on device
    if u8:
        tensors = tensors.contiguous()
    elif tensors.dtype != torch.float32 or not tensors.is_contiguous():
        tensors = tensors.to(torch.float32).contiguous()
    K = tensors.shape[0]
    dev = tensors.device
    # one coordinate set shared by all images: (2, n_s); or one set per image (unaligned patches): (K, 2, n_s)
    smp = [torch.as_tensor(np.asarray(s), dtype=torch.float64).to(dev) for s in samples]
    smp = [(s.reshape(1, 2, -1) if s.dim() == 2 else s).contiguous() for s in smp]
    for s in smp:
        if s.dim() != 3 or s.shape[1] != 2 or s.shape[0] not in (1, K):
            raise ValueError("samples[s] must be (2, n_s) or (K, 2, n_s)")
    N = int(sum(s.shape[-1] for s in smp))
    use_scales = (len(smp) > 1) if with_scales is None else with_scales
    patches = torch.zeros(K, N, 3, patch_dim, patch_dim, dtype=torch.float32, device=dev)
    pos = torch.zeros(K, N, 2, dtype=torch.float32, device=dev)
    scales = torch.zeros(K, N, dtype=torch.float32, device=dev) if use_scales else None
    levels = _u8_levels(ctx, tensors, len(smp)) if u8 else _pyramid(ctx, tensors, len(smp))
    check_coordinates(ctx, sync=False)
    off = 0
    for s, (lvl, sm) in enumerate(zip(levels, smp)):
        n = sm.shape[-1]
        if lvl is None:
            H, W = tensors.shape[1], tensors.shape[2]
            ctx.call("vtq_patch_gather_u8", _ptr(tensors), K, H, W, _ptr(sm), sm.shape[0], n, off, N, _ptr(patches),
                     None, VTQ_F16, _ptr(pos), _ptr(scales), s, _stream(dev))
        else:
            H, W = lvl.shape[-2:]
            ctx.call("vtq_patch_gather", _ptr(lvl), K, H, W, _ptr(sm), sm.shape[0], n, off, N, _ptr(patches), None,
                     VTQ_F16, _ptr(pos), _ptr(scales), s, _stream(dev))
        off += n
    check_coordinates(ctx, sync=True)   # reference-format API: raise for THIS call, like the reference does
    return patches, pos, (scales.to(torch.int32) if use_scales else None)


def get_iqa_patches(imgs, tensors, patch_count, patch_dim, patch_sampler, patch_num_scales,
                    scale_num_samples_ratio=DEFAULT_NUM_SAMPLES_RATIO, use_aligned_patches=True,
                    randomize_patch_scale_order=False, random_seed=None, debug=False):
    """Same call signature and outputs as the reference function (patch_sampling.py:450-613).  ``patch_sampler`` is
    the reference's own ``PatchSampler`` (or any object with ``compute_diff`` / ``get_sample_params``): it still
    chooses the coordinates on the host with numpy's RNG, consumed in the reference's order (slot permutation first,
    :506-508; then per level one draw, or one draw per image when ``use_aligned_patches=False``, :561); the
    difference-weighted modes get their weight map mean-pooled per level exactly like :603-605.  Extraction, uv and
    scale ids happen on the GPU (one coordinate set per image when patches are not aligned); the reference's random
    slot order (``randomize_patch_scale_order``) is a device-side permutation of the scale-ordered result."""
    if debug:
        raise NotImplementedError("vtamiq_b200.get_iqa_patches: debug=True changes the meaning of pos / scales "
                                  "(patch_sampling.py:569-583) and is not supported")
    if len(imgs) != len(tensors):
        raise ValueError("get_iqa_patches(): Image and Tensor counts should match.")
    if patch_count < patch_num_scales:
        raise ValueError("get_iqa_patches(): number of patches larger than the number of scales.")
    state = None
    if random_seed is not None:
        state = np.random.get_state()
        np.random.seed(random_seed)
    try:
        ref = imgs[0]
        height, width = (ref.height, ref.width) if hasattr(ref, "height") else ref.shape[:2]
        patch_indices = np.random.permutation(patch_count) if randomize_patch_scale_order else None
        diff = patch_sampler.compute_diff(imgs)
        n_scales = compute_patch_num_scales(patch_num_scales, height, width, patch_dim, patch_dim)
        counts = compute_num_patches_per_scale(patch_count, n_scales, scale_num_samples_ratio)
        t = torch.stack(list(tensors), dim=0)
        n_sets = 1 if use_aligned_patches else len(imgs)
        samples, h, w, total = [], t.shape[-2], t.shape[-1], 0
        for s in range(n_scales):
            n_s = int(counts[-s - 1])
            draws = [patch_sampler.get_sample_params(h, w, patch_dim, patch_dim, diff=diff, num_samples=n_s)
                     for _ in range(n_sets)]
            samples.append(np.stack([np.asarray(d, dtype=np.float64) for d in draws]))   # (n_sets, 2, n_s)
            h, w = h // 2, w // 2
            if diff is not None:   # the weight map follows the image pyramid (host side, like the sampler itself)
                d = torch.as_tensor(diff)
                diff = torch.nn.functional.avg_pool2d(d.view(1, 1, *d.shape), 2).squeeze().numpy()
            total += n_s
            if patch_count <= total:
                break
    finally:
        if state is not None:
            np.random.set_state(state)
    patches, pos, scales = extract_patches(t, samples, patch_dim, with_scales=n_scales > 1)
    if total < patch_count:   # the pyramid ran out of levels: the reference leaves the remaining slots zero
        pad = patch_count - total
        patches = torch.cat([patches, patches.new_zeros(patches.shape[0], pad, *patches.shape[2:])], dim=1)
        pos = torch.cat([pos, pos.new_zeros(pos.shape[0], pad, 2)], dim=1)
        if scales is not None:
            scales = torch.cat([scales, scales.new_zeros(scales.shape[0], pad)], dim=1)
    if patch_indices is not None:
        # slot patch_indices[i] receives the i-th patch of the scale-ordered sequence (:587-590)
        perm = torch.as_tensor(patch_indices[:patches.shape[1]], dtype=torch.long, device=patches.device)
        patches = torch.empty_like(patches).index_copy_(1, perm, patches)
        pos = torch.empty_like(pos).index_copy_(1, perm, pos)
        if scales is not None:
            scales = torch.empty_like(scales).index_copy_(1, perm, scales)
    return patches, pos, scales
