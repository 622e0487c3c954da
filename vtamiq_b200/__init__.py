"""vtamiq_b200 — B200 (sm_100a) implementation of the VTAMIQ inference hot path.

Public surface (mirrors the reference's names for this path):
  VTAMIQ, VisionTransformerBackbone ........ drop-in nn.Modules (vtamiq.py)
  get_iqa_patches, extract_patches ......... device-side patch extraction (patch_sampling.py)
  shard_pairs, gather_scores ............... batch sharding across GPUs (parallel.py)
"""
from .modules import VIT_VARIANT_B8, VIT_VARIANT_B16, VIT_VARIANT_L16, get_vit_config  # noqa: F401
from .vtamiq import VTAMIQ, VisionTransformerBackbone  # noqa: F401
from .patch_sampling import (check_coordinates, compute_num_patches_per_scale, compute_patch_num_scales,  # noqa: F401
                             extract_patches, extract_patches_batch, get_iqa_patches, perturbed_grid_samples,
                             sample_batch)
from .parallel import gather_scores, shard_pairs  # noqa: F401
from .metrics import compute_correlations  # noqa: F401

__all__ = ["VTAMIQ", "VisionTransformerBackbone", "get_iqa_patches", "extract_patches", "extract_patches_batch",
           "check_coordinates",
           "compute_patch_num_scales", "compute_num_patches_per_scale", "shard_pairs", "gather_scores", "compute_correlations", "sample_batch", "perturbed_grid_samples",
           "get_vit_config", "VIT_VARIANT_B8", "VIT_VARIANT_B16", "VIT_VARIANT_L16"]
