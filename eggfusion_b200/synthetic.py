"""Synthetic surfel scenes and cameras (numpy only, deterministic).

Implements the "layers" scene of SURVEY.md section 8(d): P/L surfels on each of L smooth depth sheets seen by a
pinhole camera, the way EGG-Fusion seeds surfels from a depth frame
(/root/reference/src/core/mapper.py:446-492, /root/reference/src/core/gaussian_surfels.py:169-222).
Camera conventions follow /root/reference/src/utils/frame.py:159-169 (viewmatrix = W2C^T,
projmatrix = (P @ W2C)^T) and /root/reference/src/utils/camera_utils.py:100-120 (getProjectionMatrix_v2).
The same arrays feed the reference, the CPU oracle and the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SEED = 20251201
SH_C0 = 0.28209479177387814


@dataclass
class Camera:
    """Everything GaussianRasterizationSettings needs, as numpy / python scalars."""

    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float
    w2c: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def fovx(self) -> float:
        return 2.0 * math.atan(self.width / (2.0 * self.fx))

    @property
    def fovy(self) -> float:
        return 2.0 * math.atan(self.height / (2.0 * self.fy))

    @property
    def tanfovx(self) -> float:
        return math.tan(self.fovx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.fovy * 0.5)

    def projection(self) -> np.ndarray:
        """getProjectionMatrix_v2 (camera_utils.py:100-120), fp32 like the reference's torch.zeros(4, 4)."""
        ty, tx = math.tan(self.fovy / 2), math.tan(self.fovx / 2)
        top, right = ty * self.znear, tx * self.znear
        P = np.zeros((4, 4), dtype=np.float32)
        P[0, 0] = 2.0 * self.znear / (2 * right)
        P[1, 1] = 2.0 * self.znear / (2 * top)
        P[3, 2] = 1.0
        P[2, 2] = self.zfar / (self.zfar - self.znear)
        P[2, 3] = -(self.zfar * self.znear) / (self.zfar - self.znear)
        return P

    @property
    def viewmatrix(self) -> np.ndarray:
        """world_view_transform = W2C^T (frame.py:159-161), contiguous fp32 [4,4]."""
        return np.ascontiguousarray(self.w2c.astype(np.float32).T)

    @property
    def projmatrix(self) -> np.ndarray:
        """full_proj_transform = W2C^T @ P^T (frame.py:163-165, dataset.py:37-44)."""
        return np.ascontiguousarray((self.viewmatrix @ self.projection().T).astype(np.float32))

    @property
    def campos(self) -> np.ndarray:
        """camera_center = inverse(W2C^T)[3, :3] (frame.py:167-169)."""
        return np.ascontiguousarray(np.linalg.inv(self.viewmatrix.astype(np.float64))[3, :3].astype(np.float32))

    @property
    def tiles(self) -> tuple[int, int]:
        return ((self.height + 15) // 16, (self.width + 15) // 16)


def default_camera(width: int, height: int, w2c: np.ndarray | None = None) -> Camera:
    """fx = fy = 0.9 W, principal point at the image centre."""
    f = 0.9 * width
    cam = Camera(width, height, f, f, (width - 1) / 2.0, (height - 1) / 2.0)
    if w2c is not None:
        cam.w2c = np.asarray(w2c, dtype=np.float32)
    return cam


def look_from(offset_xyz=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0) -> np.ndarray:
    """A W2C matrix for a camera displaced by `offset_xyz` and rotated by small yaw/pitch (radians)."""
    cy_, sy_ = math.cos(yaw), math.sin(yaw)
    cp_, sp_ = math.cos(pitch), math.sin(pitch)
    Ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
    Rx = np.array([[1, 0, 0], [0, cp_, -sp_], [0, sp_, cp_]])
    R = Rx @ Ry
    c2w = np.eye(4)
    c2w[:3, :3] = R.T
    c2w[:3, 3] = np.asarray(offset_xyz, dtype=np.float64)
    return np.linalg.inv(c2w).astype(np.float32)


def _quat_z_to(n: np.ndarray) -> np.ndarray:
    """Unit quaternion (w,x,y,z) rotating +z onto unit vectors n [P,3]
    (compute_rot / quaternion_from_axis_angle, /root/reference/src/core/utils.py:114-127)."""
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(np.broadcast_to(z, n.shape), n)
    axis = axis / (np.linalg.norm(axis, axis=-1, keepdims=True) + 1e-8)
    ang = np.arccos(np.clip(n[:, 2], -1.0, 1.0))[:, None]
    q = np.concatenate([np.cos(ang / 2), axis * np.sin(ang / 2)], axis=1)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return q.astype(np.float32)


def make_scene(P: int, cam: Camera, layers: int = 4, sh_degree: int = 3, seed: int = SEED,
               normal_jitter: float = 0.15) -> dict[str, np.ndarray]:
    """Surfels on `layers` depth sheets in the frame of a canonical (identity-pose) camera with cam's intrinsics.

    Returns post-activation tensors exactly as Renderer.render feeds them to the rasterizer
    (/root/reference/src/core/mapper.py:565-585): xyz [P,3], opacity [P,1], shs [P,M,3],
    scales [P,3] (z = 0), rotations [P,4] (w,x,y,z, unit).
    """
    rng = np.random.default_rng(seed)
    W, H = cam.width, cam.height
    M = (sh_degree + 1) ** 2
    layer = rng.integers(0, layers, size=P)
    u = rng.uniform(-0.05 * W, 1.05 * W, size=P)
    v = rng.uniform(-0.05 * H, 1.05 * H, size=P)
    un, vn = u / W, v / H
    z = 1.5 + 0.6 * layer + 0.3 * np.sin(3 * un) * np.cos(2 * vn)
    x = (u - cam.cx) * z / cam.fx
    y = (v - cam.cy) * z / cam.fy
    xyz = np.stack([x, y, z], axis=1)

    # sheet normal from the analytic surface gradient, jittered, flipped towards the camera
    dz_du = 0.3 * 3 * np.cos(3 * un) * np.cos(2 * vn) / W
    dz_dv = -0.3 * 2 * np.sin(3 * un) * np.sin(2 * vn) / H
    tu = np.stack([z / cam.fx + (u - cam.cx) / cam.fx * dz_du, (v - cam.cy) / cam.fy * dz_du, dz_du], axis=1)
    tv = np.stack([(u - cam.cx) / cam.fx * dz_dv, z / cam.fy + (v - cam.cy) / cam.fy * dz_dv, dz_dv], axis=1)
    n = np.cross(tu, tv)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n += rng.normal(0.0, normal_jitter, size=n.shape)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    flip = np.sum(n * xyz, axis=1) > 0
    n[flip] *= -1.0
    rotations = _quat_z_to(n)

    r_px = rng.uniform(1.0, 3.0, size=P)
    scales = np.stack([r_px * z / cam.fx, r_px * z / cam.fy, np.zeros(P)], axis=1)
    opacity = rng.uniform(0.3, 0.99, size=(P, 1))
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0.0, 1.0, size=(P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0.0, 0.05, size=(P, M - 1, 3))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return {"xyz": f32(xyz), "opacity": f32(opacity), "shs": f32(shs), "scales": f32(scales),
            "rotations": f32(rotations)}


def make_pixel_grads(cam: Camera, seed: int = SEED + 1, with_opacity: bool = False) -> dict[str, np.ndarray]:
    """Upstream image gradients ~ N(0,1)/N_px; dL/dopacity is zero unless requested, like
    the loss of /root/reference/src/core/mapper.py:381-438 which never touches the opacity image."""
    rng = np.random.default_rng(seed)
    H, W = cam.height, cam.width
    n = float(H * W)
    g = {
        "color": rng.normal(size=(3, H, W)) / n,
        "normal": rng.normal(size=(3, H, W)) / n,
        "depth": rng.normal(size=(1, H, W)) / n,
        "opacity": (rng.normal(size=(1, H, W)) / n) if with_opacity else np.zeros((1, H, W)),
    }
    return {k: np.ascontiguousarray(a, dtype=np.float32) for k, a in g.items()}


CONFIGS = {
    # name: (P, W, H, layers, sh_degree)   -- BASELINE.json configs / SURVEY 8 table
    "C1": (10_000, 256, 256, 2, 3),
    "C2": (600_000, 1200, 680, 4, 3),
    "C3": (1_000_000, 1920, 1080, 4, 3),
    "C4": (4_000_000, 1920, 1080, 4, 3),
    "C5": (300_000, 640, 480, 4, 0),
}


def make_config(name: str, seed: int = SEED):
    P, W, H, L, deg = CONFIGS[name]
    cam = default_camera(W, H)
    return cam, make_scene(P, cam, layers=L, sh_degree=deg, seed=seed)


def make_fusion_case(P: int, cam: Camera, seed: int = SEED + 7, alpha_p: float = 1.0, alpha_n: float = 0.5):
    """Inputs of the once-per-frame fusion kernels (`project_surfels_to_frame`, `preprocess_surfels`):
    a noisy copy of a one-sheet surfel model plus the frame's world-space vertex / normal / depth maps and masks
    (what /root/reference/src/core/mapper.py:242-308 passes).  The camera pose is cam.w2c; maps are [H,W,C]."""
    rng = np.random.default_rng(seed)
    W, H = cam.width, cam.height
    c2w = np.linalg.inv(cam.w2c.astype(np.float64))
    # analytic depth sheet in the camera frame
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    depth_of = lambda uu, vv: 1.6 + 0.25 * np.sin(3 * uu / W) * np.cos(2 * vv / H)
    z = depth_of(u, v)
    vert_c = np.stack([(u - cam.cx) * z / cam.fx, (v - cam.cy) * z / cam.fy, z], axis=-1)
    du = np.gradient(vert_c, axis=1)
    dv = np.gradient(vert_c, axis=0)
    n_c = np.cross(du, dv)
    n_c /= np.linalg.norm(n_c, axis=-1, keepdims=True)
    n_c[np.sum(n_c * vert_c, axis=-1) > 0] *= -1.0
    vmap = vert_c @ c2w[:3, :3].T + c2w[:3, 3]
    nmap = n_c @ c2w[:3, :3].T
    mask = rng.uniform(size=(H, W)) > 0.03
    holes = rng.uniform(size=(H, W)) < 0.02
    nmap[holes] = 0.0                                    # invalid normals, as a depth sensor leaves them
    dmap = z.copy()

    # surfel model: points near the sheet (most within the fusion distance), normals near the sheet normal
    us = rng.uniform(-0.08 * W, 1.08 * W, size=P)
    vs = rng.uniform(-0.08 * H, 1.08 * H, size=P)
    zs = depth_of(us, vs) + rng.normal(0, 0.012, size=P) + (rng.uniform(size=P) < 0.1) * rng.normal(0, 0.2, size=P)
    pc = np.stack([(us - cam.cx) * zs / cam.fx, (vs - cam.cy) * zs / cam.fy, zs], axis=1)
    pw = pc @ c2w[:3, :3].T + c2w[:3, 3]
    iu = np.clip(np.rint(us), 0, W - 1).astype(int)
    iv = np.clip(np.rint(vs), 0, H - 1).astype(int)
    n = n_c[iv, iu] + rng.normal(0, 0.12, size=(P, 3))
    wide = rng.uniform(size=P) < 0.08
    n[wide] += rng.normal(0, 1.5, size=(int(wide.sum()), 3))   # some far-off normals: > 60 deg and back-facing ones
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n_w = n @ c2w[:3, :3].T
    q = _quat_z_to(n_w.astype(np.float64)) * rng.uniform(0.98, 1.02, size=(P, 1)).astype(np.float32)   # raw leaf params
    sigma2 = np.stack([(zs * alpha_p) ** 2, (zs * alpha_n) ** 2], axis=1) * rng.uniform(0.3, 1.5, size=(P, 2))
    stable = rng.uniform(size=P) < 0.8
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return {
        "points": f32(pw), "rotations": f32(q), "sigma2": f32(sigma2), "stable_mask": stable,
        "intrinsic": f32([cam.fx, cam.fy, cam.cx, cam.cy]), "frame_vmap": f32(vmap), "frame_nmap": f32(nmap),
        "frame_dmap": f32(dmap), "frame_mask": mask, "fusion_dist_thres": 0.03, "alpha_p": alpha_p, "alpha_n": alpha_n,
    }


# ---------------------------------------------------------------------------------------------------------------------
# Synthetic RGB-D sequence (BASELINE configs 2 and 5: the datasets are absent offline): a camera on a smooth
# 6-DoF trajectory looking at a textured, gently curved wall, at a dataset's calibration.  Exact depth by ray / surface
# intersection, procedural colour; what /root/reference/src/utils/dataset.py's RGBDDataset.__getitem__ returns per frame
# (timestamp, uint8 colour [H,W,3], integer depth [H,W] in 1/depth_scale m, bool mask [H,W,1], 4x4 world-to-camera pose
# relative to the first frame) so that /root/reference/src/utils/frame.py:149 `Frame.init_from_dataset` consumes it.
def _wall_z(x, y):
    return 2.0 + 0.25 * np.sin(1.3 * x + 0.4) * np.cos(1.1 * y - 0.2) + 0.08 * np.sin(3.1 * x) * np.sin(2.7 * y)


def _wall_color(x, y):
    r = 0.5 + 0.35 * np.sin(9.0 * x + 1.0) * np.cos(7.0 * y) + 0.1 * np.sin(31.0 * x + 17.0 * y)
    g = 0.5 + 0.35 * np.cos(8.0 * x - 2.0 * y) + 0.1 * np.sin(23.0 * y - 5.0 * x)
    b = 0.5 + 0.3 * np.sin(6.0 * y + 0.5) * np.sin(5.0 * x + 3.0 * y) + 0.15 * np.cos(27.0 * x)
    return np.clip(np.stack([r, g, b], axis=-1), 0.0, 1.0)


def rgbd_pose(i: int) -> np.ndarray:
    """World-to-camera pose of frame i (frame 0 = identity): ~4 mm and ~0.15 degrees per frame."""
    tx, ty, tz = 0.004 * i, 0.0015 * math.sin(0.3 * i), 0.002 * i
    c2w = np.linalg.inv(look_from((tx, ty, tz), 0.0026 * i, -0.0012 * i).astype(np.float64))
    return np.linalg.inv(c2w)


def make_rgbd_frame(i: int, width: int, height: int, fx: float, fy: float, cx: float, cy: float, depth_scale: float):
    w2c = rgbd_pose(i)
    c2w = np.linalg.inv(w2c)
    v, u = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1) @ c2w[:3, :3].T     # ray directions, world
    o = c2w[:3, 3]
    s = np.full(u.shape, 2.0)                       # ray parameter (camera-frame depth, since d.z_cam = 1)
    for _ in range(12):                             # fixed point of  o.z + s d.z = wall(o.xy + s d.xy)
        p = o + s[..., None] * d
        s = (_wall_z(p[..., 0], p[..., 1]) - o[2]) / d[..., 2]
    p = o + s[..., None] * d
    color = (np.clip(_wall_color(p[..., 0], p[..., 1]), 0, 1) * 255.0 + 0.5).astype(np.uint8)
    depth = np.clip(np.rint(s * depth_scale), 0, 65535).astype(np.uint16)
    mask = np.ones((height, width, 1), dtype=bool)
    return 0.05 * i, color, depth, mask, w2c
