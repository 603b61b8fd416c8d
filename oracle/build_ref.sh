#!/usr/bin/env bash
# Builds the UNMODIFIED reference rasterizer (diff-gaussian-surfels) for sm_100a from the sources where they lie
# under /root/reference, into oracle/_ref/ (git-ignored; it still travels to the GPU box with gpurun).
# The reference's own setup.py is used on a scratch copy (/root/reference is read-only and setup.py writes into
# its source tree); the only build tweak is a forced `#include <cstdint>` that its rasterizer_impl.h forgets.
# Nothing here copies reference sources into the repository history.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${REFERENCE_ROOT:-/root/reference}/submodules/diff-gaussian-surfels"
OUT="$HERE/_ref"
WORK="${REF_BUILD_DIR:-/tmp/dgs_ref_build}"
if [ ! -d "$SRC" ]; then
    echo "reference sources not present ($SRC); keeping any prebuilt $OUT" >&2
    exit 0
fi
if ls "$OUT"/diff_gaussian_rasterization/_C*.so >/dev/null 2>&1 && [ -z "${FORCE:-}" ]; then
    echo "oracle/_ref already built"; exit 0
fi
rm -rf "$WORK" && mkdir -p "$WORK" && cp -r "$SRC/." "$WORK/" && chmod -R u+w "$WORK"
( cd "$WORK" && NVCC_APPEND_FLAGS="-include cstdint" TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS="${MAX_JOBS:-8}" \
      python setup.py build_ext --inplace > build.log 2>&1 ) || { tail -30 "$WORK/build.log"; exit 1; }
mkdir -p "$OUT/diff_gaussian_rasterization"
cp "$WORK"/diff_gaussian_rasterization/__init__.py "$WORK"/diff_gaussian_rasterization/_C*.so "$OUT/diff_gaussian_rasterization/"
echo "built $OUT"
