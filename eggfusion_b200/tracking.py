"""Dense-tracker Gauss-Newton loop on the device (C ABI: include/eggtrack.h egt_gn_*).

Host-side mirror of the reference's dense tracking (/root/reference/src/core/tracker.py:153-169,194-251):

    for l in range(pyramid_level):
        for _ in range(pyramid_iters[l]):
            dx, converged = self.tracking_optimization(pyramid_prev, pyramid_curr, level, dense_delta, uid)
            dense_delta = update_transform(dense_delta, dx)

`tracking_optimization` keeps the reference's name, arguments and return pair, but `dx` / `converged` are device
tensors (no `.item()`), and the pose is updated in place by the solve kernel.  `DenseTracker.track` runs the whole
pyramid loop -- 2 launches per step, no host sync -- and returns the final pose and the "any step converged" flag as
device tensors.  Pyramids are any objects with PyraImageCUDA's attribute lists (src/utils/frame.py:20-99).
There is no CPU fallback: host tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch

from . import _lib
from . import rasterizer as R


class TrackingConfig(NamedTuple):
    """cfg.Tracking.* (configs/replica/base.yaml:29-37)."""
    pyramid_level: int = 3
    pyramid_iters: tuple = (3, 3, 3)
    angle_threshold: float = 20.0
    distance_threshold: float = 0.1
    use_rgb: bool = True
    rgb_weight: float = 1e-4
    residual_thres: float = 0.01
    dx_threshold: float = 0.001
    lm: float = 1e-6


def _f32(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the tracker has no CPU fallback)")
    return t.float().contiguous() if t.dtype != torch.float32 or not t.is_contiguous() else t


def _u8(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the tracker has no CPU fallback)")
    t = t.contiguous()
    return t.view(torch.uint8) if t.dtype == torch.bool else (t != 0).view(torch.uint8)


def make_level(model, frame, level: int):
    """egt_level for one pyramid level + the tensors that must stay alive while it is used."""
    intr = model.intrinsic_pyramid[level]
    fx, fy, cx, cy = [float(v) for v in intr]   # a 4-element host tensor in the reference (frame.py:67,73-74)
    keep = [_f32(model.disp_pyramid[level], "model disp"), _f32(model.vertex_pyramid[level], "model vertex"),
            _f32(model.normal_pyramid[level], "model normal"), _u8(model.mask_pyramid[level], "model mask"),
            _f32(model.intensity_pyramid[level], "model intensity"), _f32(frame.vertex_pyramid[level], "frame vertex"),
            _f32(frame.normal_pyramid[level], "frame normal"), _u8(frame.mask_pyramid[level], "frame mask"),
            _f32(frame.intensity_pyramid[level], "frame intensity"), _f32(frame.grad_pyramid[level], "frame grad")]
    H, W = keep[1].shape[0], keep[1].shape[1]
    lv = _lib.Level(W, H, fx, fy, cx, cy, *[t.data_ptr() for t in keep])
    return lv, keep


class DenseTracker:
    def __init__(self, cfg: TrackingConfig = TrackingConfig(), device="cuda:0"):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        dev = self.device
        self.sums = torch.zeros(_lib.EGT_GN_SUMS, dtype=torch.float64, device=dev)
        self.dx = torch.zeros(6, dtype=torch.float32, device=dev)
        self.system = torch.zeros(42, dtype=torch.float32, device=dev)
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)

    def tracking_optimization(self, model, frame, level: int, transform: torch.Tensor, fid=None, _lv=None):
        """One Gauss-Newton step (tracker.py:194-251).  `transform` (CUDA float32 [4,4], contiguous) is updated IN PLACE
        (update_transform); returns (dx [6], converged [] bool) as device tensors."""
        cfg = self.cfg
        if not (transform.is_cuda and transform.dtype == torch.float32 and transform.is_contiguous()):
            raise RuntimeError("transform: expected a contiguous float32 CUDA tensor [4, 4]")
        lv, keep = _lv if _lv is not None else make_level(model, frame, level)
        stream = R._stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.egt_gn_accumulate(C.byref(lv), transform.data_ptr(), cfg.angle_threshold,
                                                  cfg.distance_threshold, int(cfg.use_rgb), self.sums.data_ptr(),
                                                  stream), "gn_accumulate")
            _lib.check(self.lib.egt_gn_solve_update(self.sums.data_ptr(), cfg.rgb_weight, cfg.lm, cfg.residual_thres,
                                                    cfg.dx_threshold, transform.data_ptr(), self.dx.data_ptr(),
                                                    self.system.data_ptr(), self.status.data_ptr(), stream),
                       "gn_solve_update")
        return self.dx, self.status[1] != 0

    def track(self, model, frame, delta_transform: torch.Tensor, prev_transform: torch.Tensor):
        """The dense part of Tracker.tracking_frame (tracker.py:153-169): returns (curr_transform [4,4],
        dense_converged [] bool), both on the device; nothing is read back."""
        cfg = self.cfg
        dense_delta = delta_transform.detach().clone().float().contiguous()
        made = [make_level(model, frame, level) for level in range(cfg.pyramid_level)]
        levels = (_lib.Level * cfg.pyramid_level)(*[m[0] for m in made])
        iters = (C.c_int32 * cfg.pyramid_level)(*[int(v) for v in cfg.pyramid_iters[:cfg.pyramid_level]])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.egt_track_pyramid(levels, cfg.pyramid_level, iters, cfg.angle_threshold,
                                                  cfg.distance_threshold, int(cfg.use_rgb), cfg.rgb_weight, cfg.lm,
                                                  cfg.residual_thres, cfg.dx_threshold, dense_delta.data_ptr(),
                                                  self.sums.data_ptr(), self.dx.data_ptr(), self.system.data_ptr(),
                                                  self.status.data_ptr(), R._stream_ptr(self.device)), "track_pyramid")
        conv = self.status[0] != 0
        self.last_dense_delta = dense_delta   # the optimised delta, whether or not it is committed
        curr = torch.where(conv, dense_delta @ prev_transform, delta_transform @ prev_transform)
        return curr, conv
