"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what include/eggsplat.h
declares; the Python mirror of the reference API has the reference's names and error behaviour; the product path
refuses to run without CUDA (no silent fallback)."""
import os
import re
import subprocess

import pytest
import torch

from util import ROOT

HEADERS = [os.path.join(ROOT, "include", "eggsplat.h"), os.path.join(ROOT, "include", "eggtrack.h"),
           os.path.join(ROOT, "include", "eggmap.h")]


def declared_symbols():
    src = "".join(open(h).read() for h in HEADERS)
    return sorted(set(re.findall(r"EGS_API\s+[\w\s\*]+?\b(eg[stm]_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import eggfusion_b200
    so = eggfusion_b200.build()
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT\s+(eg[stm]_\w+)", out)))
    decl = declared_symbols()
    assert len(decl) >= 17
    assert exported == decl


def test_ctypes_signatures_cover_header():
    from eggfusion_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.egs_abi_version() == _lib.ABI_VERSION
    assert b"bad argument" in lib.egs_error_string(-1)


def test_workspace_sizes_no_gpu_needed():
    from eggfusion_b200 import rasterizer as R
    g, i, b = R.workspace_sizes(1000, 256, 256, 5000)
    assert g >= 1000 * (64 + 24 + 4 + 1)
    assert i >= 256 * 256 * 12
    assert b >= 5000 * 12
    g0, i0, b0 = R.workspace_sizes(0, 16, 16, 0)
    assert g0 > 0 and i0 > 0 and b0 > 0


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """Argument checks of the sharded / fused entry points return EGS_E_BADARG / EGS_E_UNSUPPORTED before any CUDA
    call is made (so they can be exercised without a GPU): the error behaviour a binding in another language relies on."""
    import ctypes as C
    from eggfusion_b200 import _lib
    lib = _lib.load()
    BAD, UNSUP = -1, -2
    dummy = C.create_string_buffer(256)
    ptr = C.addressof(dummy)                              # a non-null pointer that is never dereferenced
    good = _lib.Frame(1000, 64, 64, 3, 16, 1.0, 1.0, 32.0, 32.0, 1.0, ptr, ptr, ptr, ptr)
    nocam = _lib.Frame(1000, 64, 64, 3, 16, 1.0, 1.0, 32.0, 32.0, 1.0, None, ptr, ptr, ptr)
    deg9 = _lib.Frame(1000, 64, 64, 9, 16, 1.0, 1.0, 32.0, 32.0, 1.0, ptr, ptr, ptr, ptr)
    sh4 = _lib.Frame(1000, 64, 64, 1, 4, 1.0, 1.0, 32.0, 32.0, 1.0, ptr, ptr, ptr, ptr)
    # sharded plan: null frame, frame without camera tensors, unsupported degree, owned range outside [0, P]
    args = [ptr] * 6 + [None]
    assert lib.egs_forward_plan_sharded(None, *args, 0, 10, ptr, ptr, ptr, ptr, None, None) == BAD
    assert lib.egs_forward_plan_sharded(C.byref(nocam), *args, 0, 10, ptr, ptr, ptr, ptr, None, None) == BAD
    assert lib.egs_forward_plan_sharded(C.byref(deg9), *args, 0, 10, ptr, ptr, ptr, ptr, None, None) == UNSUP
    assert lib.egs_forward_plan_sharded(C.byref(good), *args, 0, 10, ptr, None, ptr, ptr, None, None) == BAD   # img workspace
    # exchange: the chunk must be a positive multiple of 256 rows, the rank inside the world, the tables present
    assert lib.egs_push_rows(1000, 100, 2, 0, ptr, ptr, ptr, ptr, ptr, None) == BAD
    assert lib.egs_push_rows(1000, 512, 2, 2, ptr, ptr, ptr, ptr, ptr, None) == BAD
    assert lib.egs_push_rows(1000, 512, 2, 0, ptr, ptr, None, ptr, ptr, None) == BAD
    assert lib.egs_push_rows(1000, 512, 2, 0, None, ptr, ptr, ptr, ptr, None) == BAD
    assert lib.egs_fold_inbox(0, 2, 0, ptr, ptr, ptr, None) == BAD
    assert lib.egs_fold_inbox(512, 2, 0, None, ptr, ptr, None) == BAD
    # fused backward + SH Adam: hyper-parameters required, step >= 1, only the 16-coefficient layout
    h = _lib.AdamHyper()
    h.beta1, h.beta2, h.eps, h.step = 0.9, 0.999, 1e-8, 1
    sig = [ptr] * 11
    assert lib.egm_backward_surfels_adam(C.byref(good), 0, 10, *sig, None, ptr, ptr, None) == BAD
    h0 = _lib.AdamHyper()
    assert lib.egm_backward_surfels_adam(C.byref(good), 0, 10, *sig, C.byref(h0), ptr, ptr, None) == BAD      # step 0
    assert lib.egm_backward_surfels_adam(C.byref(good), 990, 20, *sig, C.byref(h), ptr, ptr, None) == BAD     # range
    assert lib.egm_backward_surfels_adam(C.byref(good), 0, 10, *sig, C.byref(h), None, ptr, None) == BAD      # no state
    assert lib.egm_backward_surfels_adam(C.byref(sh4), 0, 10, *sig, C.byref(h), ptr, ptr, None) == UNSUP
    assert lib.egm_backward_surfels_adam(C.byref(good), 0, 10, *sig, C.byref(h), ptr + 4, ptr, None) == UNSUP  # alignment
    assert lib.egm_backward_surfels_adam(C.byref(good), 0, 0, *sig, C.byref(h), ptr, ptr, None) == 0           # empty range
    assert b"unsupported" in lib.egs_error_string(UNSUP)


def test_sm100a_sass_present():
    so = os.path.join(ROOT, "eggfusion_b200", "libeggsplat.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_python_api_mirrors_reference_names():
    import eggfusion_b200 as E
    fields = ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
              "sh_degree", "campos", "prefiltered", "debug", "cx", "cy")
    assert E.GaussianRasterizationSettings._fields == fields
    import sys
    sys.path.insert(0, os.path.join(ROOT, "eggfusion_b200", "dropin"))
    try:
        import diff_gaussian_rasterization as D
        for name in ("GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
                     "cpu_deep_copy_tuple", "preprocess_surfels", "project_surfels_to_frame"):
            assert hasattr(D, name)
        import cuda_tracking_ext as T
        for name in ("projective_transform_cuda", "rgb_optimization_cuda", "icp_optimization_cuda",
                     "compute_vertex_and_normal_cuda", "gaussian_filter_cuda", "bilateral_filter_cuda",
                     "gaussian_downsample_cuda", "compute_gradients_cuda", "solve_block_cuda"):
            assert hasattr(T, name)      # the nine exports of tracking.cu:952-962
    finally:
        sys.path.pop(0)


def _settings():
    import eggfusion_b200 as E
    z = torch.zeros(3)
    return E.GaussianRasterizationSettings(16, 16, 0.5, 0.5, z, 1.0, torch.eye(4), torch.eye(4), 0, z, False, False,
                                           7.5, 7.5)


def test_argument_validation_matches_reference():
    import eggfusion_b200 as E
    r = E.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    o = torch.zeros(4, 1)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, o, shs=None, colors_precomp=None, scales=torch.zeros(4, 3), rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, o, shs=torch.zeros(4, 1, 3), colors_precomp=torch.zeros(4, 3), scales=torch.zeros(4, 3),
          rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(m, o, shs=torch.zeros(4, 1, 3), scales=None, rotations=None, cov3D_precomp=None)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(m, o, shs=torch.zeros(4, 1, 3), scales=torch.zeros(4, 3), rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    """On host tensors the product must fail loudly, never compute on the CPU."""
    import eggfusion_b200 as E
    r = E.GaussianRasterizer(_settings())
    with pytest.raises(RuntimeError, match="CUDA"):
        r(torch.zeros(4, 3), torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.zeros(4, 3),
          rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match=r"\(num_points, 3\)"):
        r(torch.zeros(4, 2), torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.zeros(4, 3),
          rotations=torch.zeros(4, 4))


def test_mapping_and_tracking_mirrors_refuse_host_tensors():
    """The N1 / N3 host mirrors (eggfusion_b200.mapping, .tracking) have no CPU path either."""
    import types
    from eggfusion_b200 import mapping as M, tracking as TR
    w = M.MappingWeights()
    with pytest.raises(RuntimeError, match="CUDA"):
        M.loss_seed(torch.zeros(3, 4, 4), torch.zeros(1, 4, 4), torch.zeros(3, 4, 4), torch.zeros(4, 4, 3), None, None,
                    torch.ones(4, 4, dtype=torch.bool), None, w)
    raw = {"xyz": torch.zeros(5, 3), "features_dc": torch.zeros(5, 1, 3), "features_rest": torch.zeros(5, 15, 3),
           "scaling": torch.zeros(5, 3), "rotation": torch.ones(5, 4), "opacity": torch.zeros(5, 1)}
    with pytest.raises(RuntimeError, match="CUDA"):
        M.FrameBatchOptimizer(raw, M.LrParams(1e-5, 1e-3, 1e-5, 5e-4, 1e-4))
    pyr = types.SimpleNamespace(**{k + "_pyramid": [torch.zeros(8, 8, c)] for k, c in
                                   (("disp", 1), ("vertex", 3), ("normal", 3), ("mask", 1), ("intensity", 1), ("grad", 3))},
                                intrinsic_pyramid=[torch.tensor([8.0, 8.0, 3.5, 3.5])])
    with pytest.raises(RuntimeError, match="CUDA"):
        TR.make_level(pyr, pyr, 0)
    # the ctypes structs mirror the C layouts (include/eggmap.h: egm_adam, include/eggtrack.h: egt_level)
    import ctypes
    from eggfusion_b200 import _lib
    assert ctypes.sizeof(_lib.AdamHyper) == 3 * 8 + 6 * 4 + 4 + 2 * 4 + 4   # doubles first, padded to 8
    assert ctypes.sizeof(_lib.Level) == 2 * 4 + 4 * 4 + 10 * 8


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under eggfusion_b200/ may reference it."""
    pkg = os.path.join(ROOT, "eggfusion_b200")
    for dp, _dn, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, fn
