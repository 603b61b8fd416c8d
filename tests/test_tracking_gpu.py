"""GPU parity of the dense-tracking image utilities (include/eggtrack.h) through the drop-in `cuda_tracking_ext`
module: against the CPU oracle and, when present, the goldens / live build of the reference's own kernels."""
import os
import sys

import numpy as np
import pytest
import torch

import util
from util import rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
sys.path.insert(0, os.path.join(util.ROOT, "eggfusion_b200", "dropin"))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def cuda_tracking():
    import cuda_tracking_ext as ext
    assert ext.__file__.startswith(os.path.join(util.ROOT, "eggfusion_b200"))
    ti = util.tracking_inputs()
    depth, gray, rgb = _t(ti["depth"]), _t(ti["gray"]), _t(ti["rgb"])
    H, W = depth.shape
    fx, fy, cx, cy = ti["intr"]
    out = {}
    o = torch.zeros_like(depth)
    ext.bilateral_filter_cuda(depth, o, W, H, 13, 0.03, 4.5)
    out["bilateral"] = o
    o = torch.zeros_like(rgb)
    ext.gaussian_filter_cuda(rgb, o, W, H, 3, 5, 1.5)
    out["gaussian"] = o
    for name, img in (("down1", gray[..., None].contiguous()), ("down3", rgb)):
        o = torch.zeros(H // 2, W // 2, img.shape[2], device=DEV)
        ext.gaussian_downsample_cuda(img, o, W, H, img.shape[2])
        out[name] = o
    gx, gy = torch.zeros_like(gray), torch.zeros_like(gray)
    ext.compute_gradients_cuda(gray, gx, gy, W, H)
    out["grad_x"], out["grad_y"] = gx, gy
    vm, nm = torch.zeros(H, W, 3, device=DEV), torch.zeros(H, W, 3, device=DEV)
    ext.compute_vertex_and_normal_cuda(depth, fx, fy, cx, cy, vm, nm)
    out["vertex"], out["normal"] = vm, nm
    torch.cuda.synchronize()
    return ti, {k: v.cpu().numpy() for k, v in out.items()}


def _check(out, ref, tol=2e-6):
    for k in ("bilateral", "gaussian", "down1", "down3", "grad_x", "grad_y", "vertex"):
        assert out[k].shape == ref[k].shape
        assert rel_err(out[k], ref[k]) <= tol, k
    # normals: rsqrt approximations differ in the last bits; holes give exact zeros on both sides
    assert np.array_equal(np.all(out["normal"] == 0, axis=-1), np.all(ref["normal"] == 0, axis=-1))
    assert np.abs(out["normal"] - ref["normal"]).max() <= 1e-5


def test_tracking_utils_match_oracle():
    ti, out = cuda_tracking()
    fx, fy, cx, cy = ti["intr"]
    ref = {"bilateral": orc.bilateral_filter(ti["depth"], 13, 0.03, 4.5), "gaussian": orc.gaussian_filter(ti["rgb"], 5, 1.5),
           "down1": orc.gaussian_downsample(ti["gray"][..., None]), "down3": orc.gaussian_downsample(ti["rgb"])}
    ref["grad_x"], ref["grad_y"] = orc.compute_gradient(ti["gray"])
    ref["vertex"], ref["normal"] = orc.compute_vertex_and_normal(ti["depth"], fx, fy, cx, cy)
    _check(out, ref)
    assert (np.all(out["normal"] == 0, axis=-1)).mean() > 0.01     # the hole path is exercised


def test_tracking_utils_match_reference_golden():
    path = util.golden_path("tracking_161x119")
    if not os.path.exists(path):
        pytest.skip("tracking golden not generated (reference tracking util not built)")
    ti, out = cuda_tracking()
    _check(out, dict(np.load(path)))


def test_solve_block_on_device():
    import cuda_tracking_ext as ext
    rng = np.random.default_rng(4)
    for n in (6, 3, 12):
        J = rng.normal(size=(200, n))
        A = (J.T @ J).astype(np.float32)
        b = rng.normal(size=(n, 1)).astype(np.float32)
        x = torch.zeros(n, 1, device=DEV)
        ext.solve_block_cuda(_t(A), _t(b), 1.0e-6, x)
        want = orc.solve_block(A, b, 1.0e-6)
        assert rel_err(x.cpu().numpy().reshape(-1), want) <= 1e-4
    # singular system -> zeros, no NaN
    x = torch.ones(6, 1, device=DEV)
    ext.solve_block_cuda(torch.zeros(6, 6, device=DEV), torch.ones(6, 1, device=DEV), 0.0, x)
    assert float(x.abs().max()) == 0.0


def test_reference_python_wrappers_run_unchanged():
    """/root/reference/src/utils/cuda/__init__.py imports `cuda_tracking_ext`; its wrapper functions must work on top
    of the drop-in.  Their logic is restated here (the reference tree is not available on the GPU box)."""
    import cuda_tracking_ext as cuda_tracking
    ti = util.tracking_inputs()
    depth = _t(ti["depth"])
    ht, wd = depth.shape[:2]
    fx, fy, cx, cy = torch.tensor(ti["intr"])          # `fx, fy, cx, cy = intr` with a CPU tensor (frame.py:42)
    vertex_map = torch.zeros(ht, wd, 3, device=depth.device)
    normal_map = torch.zeros(ht, wd, 3, device=depth.device)
    cuda_tracking.compute_vertex_and_normal_cuda(depth.float(), fx, fy, cx, cy, vertex_map, normal_map)
    assert float(vertex_map[..., 2].sub(depth).abs().max()) == 0.0
    A = torch.eye(6, device=DEV) * 2
    b = torch.arange(6, device=DEV, dtype=torch.float32).reshape(6, 1)
    x = torch.zeros_like(b)
    cuda_tracking.solve_block_cuda(A.float(), b.float(), 1.0e-6, x)
    assert torch.allclose(x, b / (2 + 1e-6), atol=1e-5)
    with pytest.raises(NotImplementedError):
        cuda_tracking.icp_optimization_cuda()
