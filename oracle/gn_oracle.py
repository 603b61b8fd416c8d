"""CPU oracle (numpy) for the fused dense-tracker Gauss-Newton step (include/eggtrack.h egt_gn_*) -- TEST
INFRASTRUCTURE ONLY; the product (eggfusion_b200/) never imports it.

Restates /root/reference/src/core/optimizer.py: projective_transform (:131-180), icp_optimization (:317-377),
rgb_optimization (:278-315), update_transform (:426-441) with so3_to_SO3 (src/utils/camera_utils.py:18-28), and the
tail of Tracker.tracking_optimization (src/core/tracker.py:229-251), with F.grid_sample(align_corners=True) written
out (ATen GridSampler: unnormalise to [0, size-1]; nearest = round-half-even after the border clip; bilinear with
zero padding).
PINNED: tests/golden/gn_*.npz come from the reference's own functions run with torch on the CPU
(tests/golden/make_golden_gn.py); tests/test_gn_cpu.py checks this file against them.
"""
import math

import numpy as np

F32 = np.float32


def projective_transform(T, disp, intr):
    H, W = disp.shape[:2]
    fx, fy, cx, cy = [F32(v) for v in intr]
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    us, vs = ((xx - cx) / fx).astype(F32), ((yy - cy) / fy).astype(F32)
    Ps = np.stack([us, vs, np.ones_like(us), disp.reshape(H, W).astype(F32)], -1)
    Pt = (Ps.reshape(-1, 4) @ T.T.astype(F32)).reshape(H, W, 4)
    with np.errstate(divide="ignore", invalid="ignore"):
        ut, vt, dt = Pt[..., 0] / Pt[..., 2], Pt[..., 1] / Pt[..., 2], Pt[..., 3] / Pt[..., 2]
    O = np.zeros_like(ut)
    Jc = np.stack([dt * fx, O, -ut * dt * fx, -ut * vt * fx, (1 + ut * ut) * fx, -vt * fx,
                   O, dt * fy, -vt * dt * fy, -(1 + vt * vt) * fy, ut * vt * fy, ut * fy], -1).reshape(H, W, 2, 6)
    gx = 2 * (fx * ut + cx) / F32(W - 1) - 1
    gy = 2 * (fy * vt + cy) / F32(H - 1) - 1
    return np.stack([gx, gy], -1).astype(F32), Jc.astype(F32)


def _unnorm(g, size):
    return ((g + 1) * F32(0.5) * F32(size - 1)).astype(F32)


def sample_nearest(img, coords, border):
    H, W = img.shape[:2]
    ix, iy = _unnorm(coords[..., 0], W), _unnorm(coords[..., 1], H)
    if border:
        ix, iy = np.clip(ix, 0, W - 1), np.clip(iy, 0, H - 1)
    with np.errstate(invalid="ignore"):
        jx, jy = np.rint(ix), np.rint(iy)
    ok = (jx >= 0) & (jx < W) & (jy >= 0) & (jy < H)
    jx, jy = np.where(ok, jx, 0).astype(np.int64), np.where(ok, jy, 0).astype(np.int64)
    out = img[jy, jx]
    return np.where(ok[..., None], out, 0).astype(img.dtype)


def sample_bilinear_zeros(img, coords):
    H, W = img.shape[:2]
    ix, iy = _unnorm(coords[..., 0], W), _unnorm(coords[..., 1], H)
    with np.errstate(invalid="ignore"):
        x0, y0 = np.floor(ix), np.floor(iy)
    tx, ty = ix - x0, iy - y0
    out = np.zeros(coords.shape[:2] + (img.shape[2],), F32)
    for dx_, dy_, w in ((0, 0, (1 - tx) * (1 - ty)), (1, 0, tx * (1 - ty)), (0, 1, (1 - tx) * ty), (1, 1, tx * ty)):
        xx, yy = x0 + dx_, y0 + dy_
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        xi, yi = np.where(ok, xx, 0).astype(np.int64), np.where(ok, yy, 0).astype(np.int64)
        out += np.where(ok[..., None], img[yi, xi] * w[..., None], 0).astype(F32)
    return out


def _inmask(coords, bound):
    with np.errstate(invalid="ignore"):
        return (coords[..., 0] > -bound) & (coords[..., 0] < bound) & (coords[..., 1] > -bound) & (coords[..., 1] < bound)


def icp_terms(model, frame, T, coords, angle_thres_deg, dist_thres):
    R, t = T[:3, :3].astype(F32), T[:3, 3].astype(F32)
    vprev = model["vertex"].reshape(-1, 3) @ R.T + t
    nprev = model["normal"].reshape(-1, 3) @ R.T
    vcurr = sample_nearest(frame["vertex"], coords, True).reshape(-1, 3)
    ncurr = sample_nearest(frame["normal"], coords, True).reshape(-1, 3)
    dv = vcurr - vprev
    cn = np.cross(ncurr, nprev)
    with np.errstate(invalid="ignore"):
        dist, sine = np.sqrt((dv * dv).sum(-1)), np.sqrt((cn * cn).sum(-1))
        valid = (sine < angle_thres_deg * math.pi / 180) & (dist < dist_thres)
        w = (~np.isnan(cn).any(-1)) & _inmask(coords, 0.98).reshape(-1) & (vprev[:, 2] > 0) & valid \
            & model["mask"].reshape(-1) & frame["mask"].reshape(-1)
    r = (ncurr * dv).sum(-1)
    J = np.concatenate([ncurr, np.cross(vprev, ncurr)], 1)
    J, r = J[w].astype(np.float64), r[w].astype(np.float64)
    return J.T @ J, J.T @ r, int(w.sum())


def rgb_terms(model, frame, coords, Jc):
    w = _inmask(coords, 0.90).reshape(-1) & (frame["grad"][..., 2] > 1).reshape(-1) & model["mask"].reshape(-1)
    sI = sample_bilinear_zeros(frame["intensity"], coords)
    Ji = sample_bilinear_zeros(frame["grad"][..., :2], coords)
    mcur = sample_nearest(frame["mask"].astype(F32), coords, False).reshape(-1) > 0.8
    w = w & mcur
    J = np.einsum("hwk,hwkj->hwj", Ji, Jc).reshape(-1, 6)
    r = (model["intensity"] - sI).reshape(-1)
    J, r = J[w].astype(np.float64), r[w].astype(np.float64)
    return J.T @ J, J.T @ r, int(w.sum())


def so3_exp(w):
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], F32)
    a = F32(np.linalg.norm(w))
    if a < 1e-5:
        return (np.eye(3, dtype=F32) + W + F32(0.5) * W @ W).astype(F32)
    return (np.eye(3, dtype=F32) + (np.sin(a) / a) * W + ((1 - np.cos(a)) / (a * a)) * W @ W).astype(F32)


def update_transform(T, dx):
    T = T.copy()
    T[:3, :3] = so3_exp(dx[3:]) @ T[:3, :3]
    T[:3, 3] = dx[:3] + T[:3, 3]
    return T


def gn_step(model, frame, intr, T, angle_thres_deg, dist_thres, use_rgb, rgb_weight, lm, residual_thres, dx_thres):
    """Tracker.tracking_optimization + update_transform -> dict(A, b, dx, converged, n_icp, n_rgb, T_new)."""
    coords, Jc = projective_transform(T, model["disp"], intr)
    A, b, n_icp = icp_terms(model, frame, T, coords, angle_thres_deg, dist_thres)
    n_rgb = 0
    if use_rgb:
        A2, b2, n_rgb = rgb_terms(model, frame, coords, Jc)
        A, b = A + rgb_weight * A2, b + rgb_weight * b2
    A32, b32 = A.astype(F32), b.astype(F32)
    dx = np.linalg.solve(A32.astype(np.float64) + lm * np.eye(6), b32.astype(np.float64)).astype(F32)
    residual_est = float(np.linalg.norm(b32)) / max(1.0, (n_icp + n_rgb) ** 0.5)
    converged = residual_est < residual_thres and float(np.linalg.norm(dx)) < dx_thres
    return {"A": A32, "b": b32, "dx": dx, "converged": bool(converged), "n_icp": n_icp, "n_rgb": n_rgb,
            "T_new": update_transform(T, dx)}
