#!/usr/bin/env bash
# Packs the reference files of the hot path (and the few modules they import) from the read-only reference tree
# into ONE build artefact, the git-ignored oracle/_ref/vtamiq_reference_path.tar.gz, so that the CPU arm of bench.py
# can time the UNMODIFIED reference on the GPU box's host cores (cpu_baseline.kind = "reference").  oracle/_ref/ is
# listed in .gitignore (reference sources never enter the history or the source tree) but not in .gpurunignore (the
# archive travels to the GPU box like the built .so); oracle/reference_runner.py unpacks it into a temporary
# directory at run time.
# Third-party imports absent from the image (timm, skimage, matplotlib) come from oracle/ref_shims/ (tracked; none of
# them does inference arithmetic).       usage: scripts/vendor_reference.sh [/root/reference]
set -euo pipefail
REF="${1:-${VTAMIQ_REFERENCE:-/root/reference}}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/oracle/_ref"
[ -d "$REF" ] || { echo "vendor_reference: $REF not found (nothing to do)"; exit 0; }
FILES=(
  modules/vtamiq/vtamiq.py
  modules/VisionTransformer/__init__.py
  modules/VisionTransformer/transformer.py
  modules/VisionTransformer/backbone.py
  modules/RCAN/channel_attention.py
  modules/utils.py
  data/__init__.py
  data/patch_sampling.py
  utils/__init__.py
  utils/logging/__init__.py
  utils/logging/logger.py
  utils/misc/miscelaneous.py
  utils/misc/temporary_numpy_seed.py
)
rm -rf "$DST"
mkdir -p "$DST"
for f in "${FILES[@]}"; do
  [ -f "$REF/$f" ] || { echo "vendor_reference: missing $REF/$f" >&2; exit 1; }
done
( cd "$REF" && sha256sum "${FILES[@]}" ) > "$DST/SHA256SUMS"
tar -C "$REF" --sort=name --mtime='2020-01-01' --owner=0 --group=0 --numeric-owner -czf "$DST/vtamiq_reference_path.tar.gz" "${FILES[@]}"
echo "packed ${#FILES[@]} reference files into $DST/vtamiq_reference_path.tar.gz"
