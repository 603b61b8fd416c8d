"""ctypes front-end for the CPU oracle (oracle/splat_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by
the product package.  Mirrors the stages of the reference's Rasterizer::forward / ::backward
(/root/reference/submodules/diff-gaussian-surfels/cuda_rasterizer/rasterizer_impl.cu:212-400, :404-522) and exposes
every intermediate that parity is defined on.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Cam(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("scale_modifier", C.c_float), ("bg", C.c_float * 3), ("view", C.c_float * 16),
                ("proj", C.c_float * 16), ("campos", C.c_float * 3)]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.egso_count_instances.restype = C.c_int64
        _LIB.egso_bin.restype = C.c_int
        _LIB.egso_num_threads.restype = C.c_int
    return _LIB


def num_threads() -> int:
    return int(lib().egso_num_threads())


def set_num_threads(n: int) -> None:
    lib().egso_set_num_threads(C.c_int(n))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def make_cam(width, height, tanfovx, tanfovy, cx, cy, viewmatrix, projmatrix, campos, sh_degree, sh_coeffs,
             bg=(0.0, 0.0, 0.0), scale_modifier=1.0) -> _Cam:
    cam = _Cam()
    cam.W, cam.H, cam.sh_degree, cam.sh_coeffs = int(width), int(height), int(sh_degree), int(sh_coeffs)
    cam.tanfovx, cam.tanfovy, cam.cx, cam.cy = float(tanfovx), float(tanfovy), float(cx), float(cy)
    cam.scale_modifier = float(scale_modifier)
    cam.bg[:] = [float(b) for b in bg]
    cam.view[:] = [float(v) for v in np.asarray(viewmatrix, dtype=np.float32).reshape(-1)]
    cam.proj[:] = [float(v) for v in np.asarray(projmatrix, dtype=np.float32).reshape(-1)]
    cam.campos[:] = [float(v) for v in np.asarray(campos, dtype=np.float32).reshape(-1)]
    return cam


def cam_from_synthetic(scam, sh_degree, sh_coeffs, bg=(0.0, 0.0, 0.0), scale_modifier=1.0) -> _Cam:
    """From eggfusion_b200.synthetic.Camera; tan(fov/2) is rounded through python float like render.py:55-56."""
    return make_cam(scam.width, scam.height, scam.tanfovx, scam.tanfovy, scam.cx, scam.cy, scam.viewmatrix,
                    scam.projmatrix, scam.campos, sh_degree, sh_coeffs, bg, scale_modifier)


def forward(cam: _Cam, means, scales, rotations, opacities, shs=None, colors_precomp=None, tile_mask=None,
            render: bool = True) -> dict:
    """Full forward.  Returns images plus every intermediate (geometry state, binning, per-pixel saved state)."""
    L = lib()
    means, scales, rotations = _f32(means), _f32(scales), _f32(rotations)
    opacities = _f32(opacities).reshape(-1)
    shs, colors_precomp = _f32(shs), _f32(colors_precomp)
    P = means.shape[0]
    W, H = cam.W, cam.H
    gy, gx = (H + 15) // 16, (W + 15) // 16
    tiles = gx * gy
    if tile_mask is None:
        tile_mask = np.ones((gy, gx), dtype=np.int32)
    tile_mask = np.ascontiguousarray(tile_mask, dtype=np.int32)
    o = {
        "radii": np.zeros(P, np.int32), "active_mask": np.zeros(P, np.uint8), "means2D": np.zeros((P, 2), np.float32),
        "depths": np.zeros(P, np.float32), "cov3D": np.zeros((P, 6), np.float32),
        "conic_opacity": np.zeros((P, 4), np.float32), "rgb": np.zeros((P, 3), np.float32),
        "normal": np.zeros((P, 3), np.float32), "Jinv": np.zeros((P, 10), np.float32),
        "viewCos": np.zeros(P, np.float32), "clamped": np.zeros((P, 3), np.uint8),
        "tiles_touched": np.zeros(P, np.uint32),
    }
    if P:
        L.egso_preprocess(C.byref(cam), C.c_int(P), _p(means), _p(scales), _p(rotations), _p(opacities), _p(shs),
                          _p(colors_precomp), _p(tile_mask), _p(o["radii"]), _p(o["active_mask"]), _p(o["means2D"]),
                          _p(o["depths"]), _p(o["cov3D"]), _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["normal"]),
                          _p(o["Jinv"]), _p(o["viewCos"]), _p(o["clamped"]), _p(o["tiles_touched"]))
    I = int(L.egso_count_instances(C.c_int(P), _p(o["tiles_touched"]))) if P else 0
    o["num_rendered"] = I
    o["point_list_keys"] = np.zeros(I, np.uint64)
    o["point_list"] = np.zeros(I, np.uint32)
    o["ranges"] = np.zeros((tiles, 2), np.uint32)
    o["tile_indices"] = np.full(tiles, -1, np.int32)
    o["tile_num"] = 0
    if P:
        o["tile_num"] = int(L.egso_bin(C.byref(cam), C.c_int(P), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]),
                                       _p(o["tiles_touched"]), _p(tile_mask), C.c_int64(I), _p(o["point_list_keys"]),
                                       _p(o["point_list"]), _p(o["ranges"]), _p(o["tile_indices"])))
    o["color"] = np.zeros((3, H, W), np.float32)
    o["out_normal"] = np.zeros((3, H, W), np.float32)
    o["depth"] = np.zeros((1, H, W), np.float32)
    o["opacity"] = np.zeros((1, H, W), np.float32)
    o["final_T"] = np.zeros(H * W, np.float32)
    o["final_D"] = np.zeros(H * W, np.float32)
    o["n_contrib"] = np.zeros(H * W, np.uint32)
    if P and render:
        L.egso_render_forward(C.byref(cam), C.c_int(o["tile_num"]), _p(o["tile_indices"]), _p(o["ranges"]),
                              _p(o["point_list"]), _p(o["means2D"]), _p(o["rgb"]), _p(o["normal"]), _p(o["depths"]),
                              _p(o["conic_opacity"]), _p(o["Jinv"]), _p(o["color"]), _p(o["out_normal"]),
                              _p(o["depth"]), _p(o["opacity"]), _p(o["final_T"]), _p(o["final_D"]),
                              _p(o["n_contrib"]))
    return o


def backward(cam: _Cam, fwd: dict, means, scales, rotations, shs, dL_dcolor, dL_dnormal, dL_ddepth, dL_dopacity,
             colors_precomp=None) -> dict:
    """Full backward given the dict returned by forward().  Returns screen-space and parameter gradients."""
    L = lib()
    means, scales, rotations, shs = _f32(means), _f32(scales), _f32(rotations), _f32(shs)
    P = means.shape[0]
    M = cam.sh_coeffs
    gC, gN, gD, gO = _f32(dL_dcolor), _f32(dL_dnormal), _f32(dL_ddepth), _f32(dL_dopacity)
    g = {
        "dL_dmean2D": np.zeros((P, 3), np.float32), "dL_dconic": np.zeros((P, 4), np.float32),
        "dL_dopacity": np.zeros((P, 1), np.float32), "dL_dcolors": np.zeros((P, 3), np.float32),
        "dL_dnormal": np.zeros((P, 3), np.float32), "dL_ddepth": np.zeros((P, 1), np.float32),
        "dL_dmeans3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
        "dL_dsh": np.zeros((P, M, 3), np.float32), "dL_dscales": np.zeros((P, 3), np.float32),
        "dL_drotations": np.zeros((P, 4), np.float32),
    }
    if P == 0:
        return g
    rgb = fwd["rgb"] if colors_precomp is None else _f32(colors_precomp)
    L.egso_render_backward(C.byref(cam), C.c_int(P), C.c_int(fwd["tile_num"]), _p(fwd["tile_indices"]),
                           _p(fwd["ranges"]), _p(fwd["point_list"]), _p(fwd["means2D"]), _p(rgb), _p(fwd["normal"]),
                           _p(fwd["depths"]), _p(fwd["conic_opacity"]), _p(fwd["Jinv"]), _p(fwd["final_T"]),
                           _p(fwd["final_D"]), _p(fwd["n_contrib"]), _p(gC), _p(gN), _p(gD), _p(gO),
                           _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
                           _p(g["dL_dnormal"]), _p(g["dL_ddepth"]))
    L.egso_preprocess_backward(C.byref(cam), C.c_int(P), _p(means), _p(scales), _p(rotations),
                               _p(shs if colors_precomp is None else None), _p(fwd["radii"]), _p(fwd["cov3D"]),
                               _p(fwd["clamped"]), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolors"]),
                               _p(g["dL_dnormal"]), _p(g["dL_ddepth"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
                               _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def mark_visible(means, viewmatrix, projmatrix) -> np.ndarray:
    means = _f32(means)
    P = means.shape[0]
    out = np.zeros(P, np.uint8)
    if P:
        lib().egso_mark_visible(C.c_int(P), _p(means), _p(_f32(viewmatrix).reshape(-1)), _p(_f32(projmatrix).reshape(-1)),
                                _p(out))
    return out.astype(bool)


def project_surfels(points, rotations, stable_mask, intrinsic, viewmatrix, projmatrix, height, width):
    """Oracle of project_surfels_to_frame: returns (index_map [h,w] i32, depth_buffer [h,w] f32)."""
    points, rotations = _f32(points), _f32(rotations)
    P = points.shape[0]
    stable = np.ascontiguousarray(stable_mask, dtype=np.uint8)
    index_map = np.full((height, width), -1, np.int32)
    depth = np.full((height, width), np.inf, np.float32)
    lib().egso_project_surfels(C.c_int(P), C.c_int(height), C.c_int(width), _p(points), _p(rotations), _p(stable),
                               _p(_f32(intrinsic).reshape(-1)), _p(_f32(viewmatrix).reshape(-1)),
                               _p(_f32(projmatrix).reshape(-1)), _p(index_map), _p(depth))
    return index_map, depth


def fuse_surfels(points, rotations, sigma2, intrinsic, viewmatrix, projmatrix, frame_vmap, frame_nmap, frame_dmap,
                 frame_mask, frame_imap, fusion_dist_thres, alpha_p, alpha_n):
    """Oracle of preprocess_surfels.  Returns copies: (points, rotations, sigma2, inview_mask, surface_mask)."""
    points, rotations, sigma2 = _f32(points).copy(), _f32(rotations).copy(), _f32(sigma2).copy()
    P = points.shape[0]
    ht, wd = frame_nmap.shape[0], frame_nmap.shape[1]
    inview, surface = np.zeros(P, np.uint8), np.zeros(P, np.uint8)
    fm = np.ascontiguousarray(frame_mask, dtype=np.uint8)
    im = np.ascontiguousarray(frame_imap, dtype=np.int32)
    lib().egso_fuse_surfels(C.c_int(P), C.c_int(ht), C.c_int(wd), _p(_f32(intrinsic).reshape(-1)),
                            _p(_f32(viewmatrix).reshape(-1)), _p(_f32(projmatrix).reshape(-1)), _p(_f32(frame_vmap)),
                            _p(_f32(frame_nmap)), _p(_f32(frame_dmap)), _p(fm), _p(im), _p(points), _p(rotations),
                            _p(sigma2), _p(inview), _p(surface), C.c_float(fusion_dist_thres), C.c_float(alpha_p),
                            C.c_float(alpha_n))
    return points, rotations, sigma2, inview.astype(bool), surface.astype(bool)


# ------------------------------------------------------------------------------------------- tracking utilities
def bilateral_filter(img, window, sigma_color, sigma_space):
    img = _f32(img)
    ht, wd = img.shape[:2]
    out = np.zeros_like(img)
    lib().egto_bilateral(_p(img), _p(out), C.c_int(wd), C.c_int(ht), C.c_int(window), C.c_float(sigma_color),
                         C.c_float(sigma_space))
    return out


def gaussian_filter(img, window, sigma):
    img = _f32(img)
    ht, wd, ch = img.shape
    out = np.zeros_like(img)
    lib().egto_gaussian(_p(img), _p(out), C.c_int(wd), C.c_int(ht), C.c_int(ch), C.c_int(window), C.c_float(sigma))
    return out


def gaussian_downsample(img):
    img = _f32(img)
    ht, wd, ch = img.shape
    out = np.zeros((ht // 2, wd // 2, ch), np.float32)
    lib().egto_downsample(_p(img), _p(out), C.c_int(wd), C.c_int(ht), C.c_int(ch))
    return out


def compute_gradient(img):
    img = _f32(img)
    ht, wd = img.shape[:2]
    gx, gy = np.zeros((ht, wd), np.float32), np.zeros((ht, wd), np.float32)
    lib().egto_gradients(_p(img), _p(gx), _p(gy), C.c_int(wd), C.c_int(ht))
    return gx, gy


def compute_vertex_and_normal(depth, fx, fy, cx, cy):
    depth = _f32(depth)
    ht, wd = depth.shape[:2]
    vmap, nmap = np.zeros((ht, wd, 3), np.float32), np.zeros((ht, wd, 3), np.float32)
    lib().egto_vertex_normal(_p(depth), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(vmap), _p(nmap),
                             C.c_int(wd), C.c_int(ht))
    return vmap, nmap


def solve_block(A, b, lm):
    """solveBlock (TRK:929-950) is Eigen's colPivHouseholderQr of (A + lm I) on the CPU; Eigen is not vendored in
    the reference nor installed here, so this restates it as a float64 least-squares solve (parity unpinned for
    this one function: the reference's Eigen build could not be run)."""
    A = np.asarray(A, np.float64)
    n = A.shape[0]
    return np.linalg.lstsq(A + lm * np.eye(n), np.asarray(b, np.float64).reshape(n), rcond=None)[0].astype(np.float32)
