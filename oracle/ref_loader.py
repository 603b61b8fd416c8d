"""Loads the compiled, unmodified reference rasterizer from oracle/_ref under a private module name, and decodes
its opaque workspaces.  TEST / BASELINE INFRASTRUCTURE ONLY (tests/, bench.py --impl reference, golden maker).

The byte layouts decoded here are those of GeometryState / ImageState / BinningState::fromChunk
(/root/reference/submodules/diff-gaussian-surfels/cuda_rasterizer/rasterizer_impl.cu:159-208): 128-byte aligned
bump allocation in declaration order.  Only the arrays placed before the CUB temp storage are decoded (their
offsets do not depend on CUB).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref", "diff_gaussian_rasterization")
_mod = None


def available() -> bool:
    return os.path.isdir(REF_DIR) and any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(REF_DIR))


def load():
    """Import oracle/_ref/diff_gaussian_rasterization as module `dgs_reference` (needs torch + a CUDA runtime)."""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built: run oracle/build_ref.sh where /root/reference exists")
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location("dgs_reference", os.path.join(REF_DIR, "__init__.py"),
                                                      submodule_search_locations=[REF_DIR])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["dgs_reference"] = mod
        spec.loader.exec_module(mod)
        _mod = mod
    return _mod


def _carve(buf: np.ndarray, base_addr: int, spec):
    """spec: list of (name, dtype, count, width).  Returns dict of arrays following the obtain() rule."""
    out = {}
    addr = base_addr
    for name, dt, count, width in spec:
        addr = (addr + 127) & ~127
        off = addr - base_addr
        nbytes = np.dtype(dt).itemsize * count * width
        a = buf[off:off + nbytes].view(dt)
        out[name] = a.reshape(count, width) if width > 1 else a
        addr += nbytes
    return out


def decode_geom(geom_u8, P: int) -> dict:
    """geom_u8: torch uint8 CUDA tensor (geomBuffer)."""
    base = geom_u8.data_ptr()
    buf = geom_u8.cpu().numpy()
    f, i, u = np.float32, np.int32, np.uint32
    return _carve(buf, base, [("depths", f, P, 1), ("clamped", np.uint8, P, 3), ("internal_radii", i, P, 1),
                              ("means2D", f, P, 2), ("cov3D", f, P, 6), ("conic_opacity", f, P, 4), ("rgb", f, P, 3),
                              ("normal", f, P, 3), ("Jinv", f, P, 10), ("viewCos", f, P, 1), ("pid", i, P, 1),
                              ("pview", f, P, 3), ("tiles_touched", u, P, 1)])


def decode_img(img_u8, N: int) -> dict:
    base = img_u8.data_ptr()
    buf = img_u8.cpu().numpy()
    f, u = np.float32, np.uint32
    return _carve(buf, base, [("accum_alpha", f, N, 1), ("accum_depth", f, N, 1), ("accum_color", f, N, 3),
                              ("n_contrib", u, N, 1), ("ranges", u, N, 2)])


def decode_binning(bin_u8, R: int) -> dict:
    base = bin_u8.data_ptr()
    buf = bin_u8.cpu().numpy()
    return _carve(buf, base, [("point_list", np.uint32, R, 1), ("point_list_unsorted", np.uint32, R, 1),
                              ("point_list_keys", np.uint64, R, 1)])
