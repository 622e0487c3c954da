"""Runs the UNMODIFIED reference (packed by scripts/vendor_reference.sh into the git-ignored build artefact
oracle/_ref/vtamiq_reference_path.tar.gz, unpacked here into a temporary directory) on CPU.

TEST / BENCH INFRASTRUCTURE ONLY — never imported by vtamiq_b200.  Used by bench.py's CPU arm
(``cpu_baseline.kind = "reference"``) and by tests that cross-check the oracle port against the real thing wherever
oracle/_ref/ is present.  Import shims for timm / skimage / matplotlib: oracle/ref_shims/.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "vtamiq_reference_path.tar.gz")
SHIMS = os.path.join(HERE, "ref_shims")
_unpacked = None


def available() -> bool:
    return os.path.exists(ARCHIVE)


def _import():
    global _unpacked
    if not available():
        raise RuntimeError("oracle/_ref is absent: run scripts/vendor_reference.sh where /root/reference exists")
    if _unpacked is None:
        import atexit
        import shutil
        import tarfile
        import tempfile
        _unpacked = tempfile.mkdtemp(prefix="vtamiq_ref_")
        atexit.register(shutil.rmtree, _unpacked, ignore_errors=True)
        with tarfile.open(ARCHIVE) as tf:
            tf.extractall(_unpacked, filter="data")
    for p in (_unpacked, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from data.patch_sampling import GRID_TYPE_PERTURBED_SIMPLE, PatchSampler, get_iqa_patches  # noqa: E402
    from modules.vtamiq.vtamiq import VTAMIQ  # noqa: E402
    return VTAMIQ, PatchSampler, GRID_TYPE_PERTURBED_SIMPLE, get_iqa_patches


def build_model(vit_cfg=None, vt_kwargs=None, perturb=None, seed=0):
    """modules/vtamiq/vtamiq.py VTAMIQ, random init after torch.manual_seed(seed) (pretrained=False), eval mode."""
    import contextlib
    import io
    import torch
    VTAMIQ = _import()[0]
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = VTAMIQ(vit_config=dict(pretrained=False, **(vit_cfg or {})), **(vt_kwargs or {})).eval()
    if perturb is not None:
        perturb(m)
    return m


class FixedSampler:
    """Stands in for PatchSampler inside the reference's get_iqa_patches: hands back pre-drawn coordinates, level by
    level, so that both implementations gather the very same patch sets (compute_diff -> None is what the shipped
    GRID_TYPE_PERTURBED_SIMPLE configuration returns, data/patch_sampling.py:65-69)."""

    def __init__(self, samples):
        self.samples, self.k = list(samples), 0

    def compute_diff(self, imgs):
        return None

    def get_sample_params(self, h, w, ho, wo, diff=None, num_samples=0):
        s = self.samples[self.k]
        self.k += 1
        assert s.shape[-1] == num_samples, (s.shape, num_samples)
        return s


def gather(ref_u8, dist_u8, tensors, samples, n_scales, ratio=2.0):
    """The reference's get_iqa_patches (data/patch_sampling.py:450-613) on pre-drawn coordinates."""
    get_iqa_patches = _import()[3]
    n = int(sum(s.shape[-1] for s in samples))
    return get_iqa_patches((ref_u8, dist_u8), tensors, n, 16, FixedSampler(samples), n_scales,
                           scale_num_samples_ratio=ratio, use_aligned_patches=True)
