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
    """egt_solve_block = the reference's solveBlock (Eigen colPivHouseholderQr, fp32, column-major view; tracking.cu:929-950)
    on the device, against the restatement of that algorithm (oracle/qr_oracle.py): well-conditioned, ill-conditioned,
    non-symmetric and rank-deficient systems."""
    import cuda_tracking_ext as ext
    from oracle.qr_oracle import colpiv_householder_qr_solve as qr_solve
    rng = np.random.default_rng(4)
    eps = np.finfo(np.float32).eps
    cases = []
    for n in (6, 3, 12):
        J = rng.normal(size=(200, n))
        cases.append(((J.T @ J).astype(np.float32), rng.normal(size=n).astype(np.float32), 1.0e-6))
    for p in (2, 4, 6):                                            # ill-conditioned Gauss-Newton matrices
        J = rng.normal(size=(80, 6)) * np.array([1, 1, 10.0 ** p, 1, 10.0 ** (-p / 2), 1])
        cases.append(((J.T @ J).astype(np.float32), rng.normal(size=6).astype(np.float32), 1.0e-6))
    cases.append((rng.normal(size=(6, 6)).astype(np.float32), rng.normal(size=6).astype(np.float32), 0.0))   # non-symmetric
    B = rng.normal(size=(4, 6))
    A4 = (B.T @ B).astype(np.float32)
    cases.append((A4, (A4 @ rng.normal(size=6)).astype(np.float32), 0.0))                                    # rank 4
    for A, b, lm in cases:
        n = A.shape[0]
        x = torch.zeros(n, 1, device=DEV)
        ext.solve_block_cuda(_t(A), _t(b.reshape(n, 1)), lm, x)
        got = x.cpu().numpy().reshape(-1)
        want, rank = qr_solve(A, b, lm)
        assert np.array_equal(got == 0, want == 0), (n, rank)                     # same pivots dropped
        M = A.T.astype(np.float64) + lm * np.eye(n)
        cond = min(np.linalg.cond(M), 1e7) if rank == n else 1e4
        assert np.abs(got - want).max() <= 20 * eps * cond * np.abs(want).max() + 1e-7, (n, rank, cond)
        if rank == n and cond < 1e5:                                              # and both solve the system
            x64 = np.linalg.solve(M, b.astype(np.float64))
            assert rel_err(got, x64) <= 1e-3
    # the damped zero system the tracker can produce when no pixel is valid: (0 + lm I) x = b
    x = torch.ones(6, 1, device=DEV)
    ext.solve_block_cuda(torch.zeros(6, 6, device=DEV), torch.ones(6, 1, device=DEV), 1.0e-6, x)
    assert torch.allclose(x, torch.full_like(x, 1.0e6), rtol=1e-5)


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


# ---- fused frame ingest against the reference's own sequence --------------------------------------
def _pyramid_sequence(ext, color, depth_raw, mask, intr, nlevel=3):
    """Frame.__init__ + PyraImageCUDA._build_pyramid of the reference (src/utils/frame.py:112-146, :32-99) call for call,
    through the extension module `ext` (its own build, or the per-function drop-in): the oracle of the fused chain."""
    import torch.nn.functional as F

    def call(fn, x, *shape_out, args=()):
        o = torch.zeros(*shape_out, device=x.device)
        fn(x, o, *args)
        return o
    H, W = depth_raw.shape[:2]
    depth = call(ext.bilateral_filter_cuda, depth_raw, H, W, 1, args=(W, H, 13, 0.03, 4.5))            # frame.py:132
    gray = (color[..., 0] * 0.114 + color[..., 1] * 0.587 + color[..., 2] * 0.299)[..., None]            # frame.py:40
    fx, fy, cx, cy = intr
    vmap, nmap = torch.zeros(H, W, 3, device=DEV), torch.zeros(H, W, 3, device=DEV)
    ext.compute_vertex_and_normal_cuda(depth, fx, fy, cx, cy, vmap, nmap)                                 # frame.py:42
    P = {k: [] for k in ("gray", "disp", "mask", "vertex", "normal", "grad", "depth")}

    def grad_of(g):
        h, w = g.shape[:2]
        gx, gy = torch.zeros(h, w, device=DEV), torch.zeros(h, w, device=DEV)
        ext.compute_gradients_cuda(g.contiguous(), gx, gy, w, h)
        return torch.stack([gx, gy, torch.sqrt(gx ** 2 + gy ** 2 + 1e-6)], dim=-1).contiguous()

    def down(x):
        h, w, c = x.shape
        o = torch.zeros(h // 2, w // 2, c, device=DEV)
        ext.gaussian_downsample_cuda(x.contiguous().float(), o, w, h, c)
        return o
    m = mask
    P["gray"].append(gray); P["vertex"].append(vmap); P["normal"].append(nmap); P["depth"].append(depth)
    P["disp"].append(1.0 / (depth + 1e-6)); P["mask"].append((m > 0.9) & (depth > 0.1)); P["grad"].append(grad_of(gray))
    for l in range(1, nlevel):                                                                            # frame.py:76-99
        gray = down(gray)
        depth = down(depth)
        h, w = depth.shape[:2]
        depth = call(ext.bilateral_filter_cuda, depth, h, w, 1, args=(w, h, 13, 0.03, 4.5))
        m = down(m)
        P["gray"].append(gray); P["depth"].append(depth); P["disp"].append(1.0 / (depth + 1e-6))
        P["mask"].append((m > 0.9) & (depth > 0.1))
        P["vertex"].append(down(P["vertex"][-1]))
        P["normal"].append(F.normalize(down(P["normal"][-1]), dim=-1))
        P["grad"].append(grad_of(gray))
    return P


def _ingest_inputs(W, H, seed=3):
    rng = np.random.default_rng(seed)
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    depth = 1.6 + 0.4 * np.sin(u / 37.0) * np.cos(v / 23.0) + 0.003 * rng.normal(size=(H, W))
    depth[(u > 0.6 * W) & (v < 0.3 * H)] += 0.7                # a depth edge
    depth[rng.uniform(size=(H, W)) < 0.01] = 0.0               # sensor holes
    color = np.stack([0.5 + 0.4 * np.sin(u / 9.0 + v / 13.0), 0.5 + 0.4 * np.cos(u / 7.0), 0.5 + 0.3 * np.sin(v / 5.0)], -1)
    color += 0.02 * rng.normal(size=color.shape)
    mask = (rng.uniform(size=(H, W)) > 0.02).astype(np.float32)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return _t(f32(np.clip(color, 0, 1))), _t(f32(depth)[..., None]), _t(mask[..., None])


@pytest.mark.parametrize("W,H", [(161, 119), (640, 480), (1200, 680)])
def test_fused_ingest_matches_the_per_function_sequence(W, H):
    """egt_ingest_frame (3 launches) against the reference's call sequence through (a) the per-function drop-in kernels,
    which are pinned on the reference's own kernels, and (b) the reference's own build when oracle/_ref travelled."""
    from eggfusion_b200 import tracking as TRK
    import cuda_tracking_ext as ours_ext
    color, depth_raw, mask = _ingest_inputs(W, H)
    intr = (0.5 * W, 0.5 * W, (W - 1) / 2.0, (H - 1) / 2.0)
    pyr = TRK.FrameIngest(W, H, 3, device=DEV)(color, depth_raw, mask, intr)
    exts = [("drop-in", ours_ext)]
    mg = util._golden_mod()
    ref_ext = mg.load_ref_tracking()
    if ref_ext is not None:
        exts.append(("reference", ref_ext))
    got = {"gray": pyr.intensity_pyramid, "disp": pyr.disp_pyramid, "mask": pyr.mask_pyramid, "vertex": pyr.vertex_pyramid,
           "normal": pyr.normal_pyramid, "grad": pyr.grad_pyramid, "depth": pyr.depth_pyramid}
    for name, ext in exts:
        want = _pyramid_sequence(ext, color, depth_raw, mask, intr)
        torch.cuda.synchronize()
        for l in range(3):
            m_w, m_g = want["mask"][l], got["mask"][l]
            # the bool mask thresholds depth > 0.1 and mask > 0.9: identical except where a value sits on the threshold
            assert float((m_w != m_g).float().mean()) <= 1e-5, (name, l)
            for k in ("gray", "depth", "vertex", "grad"):
                a, b = got[k][l], want[k][l]
                assert a.shape == b.shape, (name, k, l, a.shape, b.shape)
                e = float((a - b).abs().max() / (b.abs().max() + 1e-30))
                assert e <= 2e-6, (name, k, l, e)
            # disparity of sensor holes is 1/(0 + 1e-6): compare where the depth is valid
            ok = want["depth"][l] > 0.05
            e = float(((got["disp"][l] - want["disp"][l]).abs()[ok]).max() / want["disp"][l][ok].abs().max())
            assert e <= 2e-6, (name, "disp", l, e)
            nz_w, nz_g = (want["normal"][l] == 0).all(-1), (got["normal"][l] == 0).all(-1)
            assert float((nz_w != nz_g).float().mean()) <= 1e-5, (name, "normal holes", l)
            assert float((got["normal"][l] - want["normal"][l]).abs().max()) <= 2e-5, (name, "normal", l)
    # the intrinsics list keeps the reference's quirk (level 2 = level 0 / 8, frame.py:80-81)
    assert torch.allclose(pyr.intrinsic_pyramid[2], torch.tensor(intr) / 8.0)
