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
    echo "oracle/_ref rasterizer already built"
else
rm -rf "$WORK" && mkdir -p "$WORK" && cp -r "$SRC/." "$WORK/" && chmod -R u+w "$WORK"
( cd "$WORK" && NVCC_APPEND_FLAGS="-include cstdint" TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS="${MAX_JOBS:-8}" \
      python setup.py build_ext --inplace > build.log 2>&1 ) || { tail -30 "$WORK/build.log"; exit 1; }
mkdir -p "$OUT/diff_gaussian_rasterization"
cp "$WORK"/diff_gaussian_rasterization/__init__.py "$WORK"/diff_gaussian_rasterization/_C*.so "$OUT/diff_gaussian_rasterization/"
fi
echo "built $OUT (rasterizer)"

# --- the dense-tracking util (src/utils/cuda): one .cu file; needs <Eigen/Dense> only for its host-side solveBlock.
# Eigen is not installed here, so a stub header (oracle/stubs/Eigen/Dense) lets the unmodified file compile; its CUDA
# kernels (what the goldens use) are untouched, solveBlock aborts if called.
TRK_SRC="${REFERENCE_ROOT:-/root/reference}/src/utils/cuda"
if [ -d "$TRK_SRC" ] && ! ls "$OUT"/cuda_tracking_ext*.so >/dev/null 2>&1; then
    TW="${REF_BUILD_DIR:-/tmp/dgs_ref_build}_trk"
    rm -rf "$TW" && mkdir -p "$TW" && cp -r "$TRK_SRC/." "$TW/" && chmod -R u+w "$TW"
    ( cd "$TW" && CPLUS_INCLUDE_PATH="$HERE/stubs" NVCC_APPEND_FLAGS="-I$HERE/stubs" TORCH_CUDA_ARCH_LIST="10.0a" \
          MAX_JOBS="${MAX_JOBS:-8}" python setup.py build_ext --inplace > build.log 2>&1 ) || { tail -30 "$TW/build.log"; exit 1; }
    cp "$TW"/cuda_tracking_ext*.so "$OUT/"
    echo "built $OUT (tracking util, Eigen stubbed)"
fi

# --- the reference's own Python (core loop + configs), byte for byte, so that the loop test can run it UNMODIFIED on
# the GPU box (where /root/reference does not exist) on top of either rasterizer.  Git-ignored like the rest of _ref.
EGG_SRC="${REFERENCE_ROOT:-/root/reference}"
if [ -d "$EGG_SRC/src" ] && { [ ! -d "$OUT/egg/src" ] || [ -n "${FORCE:-}" ]; }; then
    rm -rf "$OUT/egg" && mkdir -p "$OUT/egg"
    cp -r "$EGG_SRC/src" "$EGG_SRC/configs" "$EGG_SRC/main.py" "$OUT/egg/"
    find "$OUT/egg" -name "__pycache__" -type d -prune -exec rm -rf {} + 2>/dev/null || true
    rm -rf "$OUT/egg/src/utils/cuda/src" "$OUT/egg/src/utils/cuda/build" 2>/dev/null || true   # CUDA sources stay where they are
    echo "copied the reference's src/ and configs/ to $OUT/egg"
fi
