"""Surfel fusion entry points of the reference extension, on top of libeggsplat.so.

Mirrors `preprocess_surfels` and `project_surfels_to_frame` of
/root/reference/submodules/diff-gaussian-surfels/diff_gaussian_rasterization/__init__.py:233-331 (same positional
signatures, in-place semantics and return values), which /root/reference/src/core/mapper.py:268-308 calls once per
frame.  Differences: launched on the current stream without a device synchronisation, and the index map is
race-free (see egs_fusion.cu).
"""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _f32c, _stream_ptr


def _inplace_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """Tensors the kernels mutate must be written where the caller sees them: no silent copies."""
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"eggsplat: `{name}` must be a contiguous float32 CUDA tensor (it is updated in place)")
    return t


def project_surfels_to_frame(points, rotations, stable_mask, intrinsic, viewmatrix, projmatrix, factor, height, width):
    """Closest stable surfel per pixel.  Returns (index_map [h, w] int32 with -1 = no hit, depth_buf [h, w] float32
    with +inf = no hit), h = int(factor * height), w = int(factor * width) like the reference."""
    lib = _lib.load()
    ht, wd = int(factor * height), int(factor * width)
    points = points.detach()
    if not points.is_cuda:
        raise RuntimeError("eggsplat: points must be a CUDA tensor (there is no CPU implementation)")
    dev = points.device
    with torch.cuda.device(dev):
        index_map = torch.full((ht, wd), -1, dtype=torch.int32, device=dev)
        depth_buf = torch.full((ht, wd), float("inf"), dtype=torch.float32, device=dev)
        P = points.shape[0]
        pts, rot = _f32c(points, dev), _f32c(rotations.detach(), dev)
        stable = stable_mask.to(device=dev, dtype=torch.bool).contiguous()
        intr, view, proj = _f32c(intrinsic, dev), _f32c(viewmatrix, dev), _f32c(projmatrix, dev)
        scratch = torch.empty((max(ht * wd, 1),), dtype=torch.int64, device=dev)
        _lib.check(lib.egs_project_surfels(P, ht, wd, pts.data_ptr() if P else None, rot.data_ptr() if P else None,
                                           stable.data_ptr() if P else None, intr.data_ptr(), view.data_ptr(),
                                           proj.data_ptr(), scratch.data_ptr(), index_map.data_ptr(),
                                           depth_buf.data_ptr(), _stream_ptr(dev)), "project_surfels")
    return index_map, depth_buf


def preprocess_surfels(points, rotations, scales, colors, confidence, tic, eta, sigma2, observe_count, error_count,
                       stable_mask, intrinsic, viewmatrix, projmatrix, frame_vmap, frame_nmap, frame_cmap, frame_dmap,
                       frame_mask, frame_imap, depth_buff, model_vmap, model_nmap, model_mask, inview_mask,
                       surface_mask, fusion_dist_thres, alpha_p, alpha_n):
    """In-place fusion of `points`, `rotations`, `sigma2`; writes `inview_mask`, `surface_mask`.  The arguments the
    reference kernel never reads are accepted and ignored."""
    lib = _lib.load()
    pts = points.detach()
    if not pts.is_cuda:
        raise RuntimeError("eggsplat: points must be a CUDA tensor (there is no CPU implementation)")
    dev = pts.device
    P = pts.shape[0]
    if P == 0:
        return
    with torch.cuda.device(dev):
        pts = _inplace_f32(pts, "points")
        rot = _inplace_f32(rotations.detach(), "rotations")
        s2 = _inplace_f32(sigma2.detach(), "sigma2")
        for name, m in (("inview_mask", inview_mask), ("surface_mask", surface_mask)):
            if not (m.is_cuda and m.dtype == torch.bool and m.is_contiguous()):
                raise RuntimeError(f"eggsplat: `{name}` must be a contiguous bool CUDA tensor (it is written in place)")
        ht, wd = frame_nmap.shape[0], frame_nmap.shape[1]
        intr, view, proj = _f32c(intrinsic, dev), _f32c(viewmatrix, dev), _f32c(projmatrix, dev)
        vmap, nmap, dmap = _f32c(frame_vmap, dev), _f32c(frame_nmap, dev), _f32c(frame_dmap, dev)
        fmask = frame_mask.to(device=dev, dtype=torch.bool).contiguous()
        imap = frame_imap.to(device=dev, dtype=torch.int32).contiguous()
        _lib.check(lib.egs_fuse_surfels(P, ht, wd, intr.data_ptr(), view.data_ptr(), proj.data_ptr(), vmap.data_ptr(),
                                        nmap.data_ptr(), dmap.data_ptr(), fmask.data_ptr(), imap.data_ptr(),
                                        pts.data_ptr(), rot.data_ptr(), s2.data_ptr(), inview_mask.data_ptr(),
                                        surface_mask.data_ptr(), float(fusion_dist_thres), float(alpha_p),
                                        float(alpha_n), _stream_ptr(dev)), "fuse_surfels")
