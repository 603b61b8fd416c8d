"""ctypes binding of libeggsplat.so (include/eggsplat.h).  The product path has no CPU fallback: if the CUDA
library is missing or fails to load, importing a rasterizer entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# EGS_LIB: developer hook for A/B runs of alternative builds of the same library (profiles/ab_variants.sh)
LIB_PATH = os.environ.get("EGS_LIB") or os.path.join(_HERE, "libeggsplat.so")
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 1

EGS_FWD_REUSE_BINNING = 1
EGS_FWD_NO_SAVE = 2
EGS_BWD_GRADS_PREZEROED = 1
SCREEN_GRAD_STRIDE = 16


class Frame(C.Structure):
    """struct egs_frame"""
    _fields_ = [("num_surfels", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("sh_degree", C.c_int32),
                ("sh_coeffs", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("scale_modifier", C.c_float), ("bg", C.c_void_p), ("viewmatrix", C.c_void_p),
                ("projmatrix", C.c_void_p), ("campos", C.c_void_p)]


class Counters(C.Structure):
    """struct egs_counters"""
    _fields_ = [("num_rendered", C.c_int32), ("tile_num", C.c_int32), ("overflow", C.c_int32),
                ("num_visible", C.c_int32)]


# name -> (restype, argtypes); must list every symbol include/eggsplat.h declares (tests check this)
_P, _I32, _I64 = C.c_void_p, C.c_int32, C.c_int64
SIGNATURES = {
    "egs_abi_version": (C.c_int, []),
    "egs_error_string": (C.c_char_p, [C.c_int]),
    "egs_workspace_sizes": (C.c_int, [_I32, _I32, _I32, _I64, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_size_t)]),
    "egs_bin_bytes_forward_only": (C.c_int, [_I64, C.POINTER(C.c_size_t)]),
    "egs_forward_plan": (C.c_int, [C.POINTER(Frame)] + [_P] * 13),
    "egs_forward_plan_sharded": (C.c_int, [C.POINTER(Frame)] + [_P] * 7 + [_I32, _I32] + [_P] * 6),
    "egs_push_rows": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "egs_fold_inbox": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P]),
    "egs_forward_render": (C.c_int, [C.POINTER(Frame), _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I32, _P]),
    "egs_backward_render": (C.c_int, [C.POINTER(Frame), _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I32, _P]),
    "egs_backward_surfels": (C.c_int, [C.POINTER(Frame), _I32, _I32] + [_P] * 17),
    "egs_mark_visible": (C.c_int, [_I32, _P, _P, _P, _P, _P]),
    "egs_project_surfels": (C.c_int, [_I32, _I32, _I32] + [_P] * 10),
    "egs_fuse_surfels": (C.c_int, [_I32, _I32, _I32] + [_P] * 13 + [C.c_float, C.c_float, C.c_float, _P]),
    "egs_debug_export": (C.c_int, [C.POINTER(Frame), _P, _P, _P, _I64] + [_P] * 11),
}

# include/eggtrack.h
SIGNATURES.update({
    "egt_bilateral_filter": (C.c_int, [_P, _P, _I32, _I32, _I32, C.c_float, C.c_float, _P]),
    "egt_gaussian_filter": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, C.c_float, _P]),
    "egt_gaussian_downsample": (C.c_int, [_P, _P, _I32, _I32, _I32, _P]),
    "egt_compute_gradients": (C.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "egt_vertex_normal_map": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _I32, _I32, _P]),
    "egt_solve_block": (C.c_int, [_P, _P, C.c_float, _P, _I32, _P]),
})


class Level(C.Structure):
    """struct egt_level"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("model_disp", C.c_void_p), ("model_vertex", C.c_void_p),
                ("model_normal", C.c_void_p), ("model_mask", C.c_void_p), ("model_intensity", C.c_void_p),
                ("frame_vertex", C.c_void_p), ("frame_normal", C.c_void_p), ("frame_mask", C.c_void_p),
                ("frame_intensity", C.c_void_p), ("frame_grad", C.c_void_p)]


class PyramidLevel(C.Structure):
    """struct egt_pyramid_level"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("depth", C.c_void_p), ("disp", C.c_void_p),
                ("mask", C.c_void_p), ("maskf", C.c_void_p), ("vertex", C.c_void_p), ("normal", C.c_void_p),
                ("gray", C.c_void_p), ("grad", C.c_void_p)]


EGT_GN_SUMS = 56
SIGNATURES.update({
    "egt_gn_accumulate": (C.c_int, [C.POINTER(Level), _P, C.c_float, C.c_float, _I32, _P, _P]),
    "egt_gn_solve_update": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "egt_ingest_frame": (C.c_int, [_P, _P, _P, _I32, _I32] + [C.c_float] * 6 + [_I32, C.POINTER(PyramidLevel), _P]),
    "egt_track_pyramid": (C.c_int, [_P, _I32, _P, C.c_float, C.c_float, _I32, C.c_float, C.c_float, C.c_float,
                                    C.c_float, _P, _P, _P, _P, _P, _P]),
})

# include/eggmap.h
class AdamHyper(C.Structure):
    """struct egm_adam"""
    _fields_ = [("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("lr_xyz", C.c_float), ("lr_f_dc", C.c_float), ("lr_f_rest", C.c_float), ("lr_opacity", C.c_float),
                ("lr_scaling", C.c_float), ("lr_rotation", C.c_float), ("step", C.c_int32), ("reg_weight", C.c_float),
                ("reg_weight_n", C.c_float)]


EGM_TERMS, EGM_REG = 8, 4
SIGNATURES.update({
    "egm_loss_seed": (C.c_int, [_I32, _I32] + [_P] * 8 + [C.c_float] * 3 + [_P] * 5),
    "egm_loss_seed_tiles": (C.c_int, [_I32, _I32] + [_P] * 9 + [C.c_float] * 3 + [_P] * 5),
    "egm_adam_step": (C.c_int, [_I32, _I32, C.POINTER(AdamHyper)] + [_P] * 27),
    "egm_backward_surfels_adam": (C.c_int, [C.POINTER(Frame), _I32, _I32] + [_P] * 11 + [C.POINTER(AdamHyper), _P, _P, _P]),
    "egm_activate": (C.c_int, [_I32] + [_P] * 8),
    "egm_loss_total": (C.c_int, [_P, _P, _I32, _I32] + [C.c_float] * 5 + [_I32, _I32, _P, _P]),
})

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libeggsplat.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libeggsplat.so failed")
    return LIB_PATH


def load():
    """Load the CUDA library.  Raises if it is absent: there is deliberately no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C {CSRC}` (or eggfusion_b200.build()); "
            "the rasterizer has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.egs_abi_version() != ABI_VERSION:
        raise RuntimeError("libeggsplat.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().egs_error_string(rc).decode()
        raise RuntimeError(f"eggsplat {what} failed: {msg} (code {rc})")
