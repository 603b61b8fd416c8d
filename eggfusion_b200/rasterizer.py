"""Host side of the surfel rasterizer: the reference's Python extension API on top of libeggsplat.so.

Mirrors /root/reference/submodules/diff-gaussian-surfels/diff_gaussian_rasterization/__init__.py:
`GaussianRasterizationSettings` (:166-180), `GaussianRasterizer` (:182-230), `rasterize_gaussians` (:21-42) and the
autograd function `_RasterizeGaussians` (:44-164) -- same names, argument meaning, return tuples and error
behaviour -- so /root/reference/src/core/render.py:53-104 runs unchanged.  Tensor allocation, autograd wiring and
stream selection live here (PyTorch is the plumbing); all computation is in the CUDA library.  There is no CPU
path: without a CUDA device / the built library every entry point raises.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import Counters, Frame


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    cx: float
    cy: float


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


# ------------------------------------------------------------------------------------------------- helpers
def _ptr(t: Optional[torch.Tensor]):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    """What `.contiguous().data<float>()` does in the reference binding, plus moving stray host tensors over."""
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _present(t: Optional[torch.Tensor]) -> bool:
    """The reference encodes "not provided" as an empty CPU tensor (__init__.py:207-217)."""
    return t is not None and t.numel() > 0


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on `device` (the raw getter skips building a Stream object)."""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _Config:
    """Instance-capacity policy of the binning workspace.

    "exact": read the instance count back after the per-surfel stage (one 16-byte D2H + stream sync per forward;
             the reference needs two blocking copies plus a host loop, rasterizer_impl.cu:311,349-366).
    int    : fixed capacity in instances, no host sync at all; an overflow raises at the next forward/backward.
    "auto" : "exact" the first time a (surfel count, image size) is seen, afterwards `auto_headroom` x (+ `auto_slack`) the largest
             instance count reported by the last `auto_window` forwards of that shape -- no host sync in steady state
             (an optimisation loop's instance count drifts slowly).  Counts come back asynchronously; an overflow
             raises at the next forward/backward like the fixed capacity does.
    """
    capacity = "exact"
    auto_headroom = 1.5
    auto_slack = 4096       # instances added on top of the head-room (small scenes fluctuate relatively more)
    auto_window = 8


config = _Config()
_pending_overflow = []  # (event, pinned counters, capacity) of no-sync forwards not yet checked
_pinned_free = []       # recycled 4 x int32 pinned host buffers (cudaHostAlloc is far too slow to do per call)


def _get_pinned():
    return _pinned_free.pop() if _pinned_free else torch.empty((4,), dtype=torch.int32).pin_memory()


_auto_history = {}      # (device index, P, W, H) -> recent instance counts ("auto" capacity)
_captured = []          # counter words of forwards recorded into a CUDA graph (checked by check_captured())


def _capturing() -> bool:
    return torch.cuda.is_current_stream_capturing()


def check_captured(clear: bool = False):
    """Forwards recorded into a CUDA graph (torch.cuda.graph) cannot report a binning overflow themselves: nothing in
    a graph may talk to the host.  Call this after a replay (it synchronises the device): raises if any captured
    forward produced more instances than its capacity; returns the list of (num_rendered, tile_num, overflow,
    num_visible) tuples otherwise."""
    out = []
    for words, cap in _captured:
        c = tuple(int(v) for v in words.cpu())
        if c[2] != 0:
            raise RuntimeError(
                f"eggsplat: a graph-captured forward produced {c[0]} instances but the binning workspace was sized for "
                f"{cap}; its lists were truncated. Re-capture with a larger eggfusion_b200.rasterizer.config.capacity.")
        out.append(c)
    if clear:
        _captured.clear()
    return out


def _auto_note(key, count: int):
    h = _auto_history.setdefault(key, [])
    h.append(int(count))
    del h[:-config.auto_window]


def _check_pending(block: bool = False):
    while _pending_overflow:
        ev, host, cap, key = _pending_overflow[0]
        if not block and not ev.query():
            return
        ev.synchronize()
        _pending_overflow.pop(0)
        _pinned_free.append(host)
        if key is not None:
            _auto_note(key, int(host[0]))
        if int(host[2]) != 0:
            raise RuntimeError(
                f"eggsplat: a previous forward produced {int(host[0])} instances but the binning workspace was sized "
                f"for {cap}; its lists were truncated. Raise eggfusion_b200.rasterizer.config.capacity or use 'exact'.")


def make_frame(P: int, settings: GaussianRasterizationSettings, sh_coeffs: int, device):
    """egs_frame plus the (kept-alive) contiguous device copies of the four small tensors."""
    bg = _f32c(settings.bg, device)
    view = _f32c(settings.viewmatrix, device)
    proj = _f32c(settings.projmatrix, device)
    campos = _f32c(settings.campos, device)
    fr = Frame(int(P), int(settings.image_width), int(settings.image_height), int(settings.sh_degree),
               int(sh_coeffs), float(settings.tanfovx), float(settings.tanfovy), float(settings.cx),
               float(settings.cy), float(settings.scale_modifier), bg.data_ptr(), view.data_ptr(), proj.data_ptr(),
               campos.data_ptr())
    return fr, (bg, view, proj, campos)


_ws_cache = {}


def workspace_sizes(P: int, W: int, H: int, cap: int):
    key = (P, W, H, cap)
    hit = _ws_cache.get(key)
    if hit is not None:
        return hit
    g, i, b = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _lib.check(_lib.load().egs_workspace_sizes(P, W, H, cap, C.byref(g), C.byref(i), C.byref(b)), "workspace_sizes")
    if len(_ws_cache) > 256:
        _ws_cache.clear()
    _ws_cache[key] = (g.value, i.value, b.value)
    return _ws_cache[key]


class ForwardState:
    """Everything a forward leaves behind for the backward (the reference's geomBuffer / binningBuffer /
    imgBuffer / tile_indices / radii saved set, __init__.py:97-100)."""
    __slots__ = ("frame", "keep", "geom", "img", "bin", "cap", "radii", "num_rendered", "tile_num", "tile_mask", "saved")


def _bin_bytes_forward_only(cap: int) -> int:
    b = C.c_size_t()
    _lib.check(_lib.load().egs_bin_bytes_forward_only(cap, C.byref(b)), "bin_bytes_forward_only")
    return b.value


def forward_raw(settings, means3D, shs, colors_precomp, opacities, scales, rotations, tile_mask, capacity=None,
                save=True, own_range=None):
    """Run the CUDA forward.  Returns (color, normal, depth, opacity, active_mask, radii, ForwardState).
    save=False: forward-only render (EGS_FWD_NO_SAVE): nothing is kept for a backward, the binning workspace is 12
    instead of 76 bytes per instance.
    own_range=(first, count): one rank of a tile-sharded frame (parallel.py): colour / record work is skipped for
    surfels that touch none of `tile_mask`'s tiles and lie outside the owned rows (egs_forward_plan_sharded)."""
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:60-62
    if not means3D.is_cuda:
        raise RuntimeError("eggsplat: means3D must be a CUDA tensor (there is no CPU rasterizer)")
    device = means3D.device
    P = means3D.size(0)
    H, W = int(settings.image_height), int(settings.image_width)
    with torch.cuda.device(device):
        capturing = _capturing()
        if not capturing:
            _check_pending()
        stream = _stream_ptr(device)
        means3D = _f32c(means3D, device)
        use_sh = _present(shs)
        M = 0
        if use_sh:
            shs = _f32c(shs, device)
            M = shs.size(1)
        else:
            colors_precomp = _f32c(colors_precomp, device)
        opacities = _f32c(opacities, device)
        scales = _f32c(scales, device)
        rotations = _f32c(rotations, device)
        if tile_mask is not None:
            tile_mask = tile_mask.to(device=device, dtype=torch.int32).contiguous()
        frame, keep = make_frame(P, settings, M, device)

        u8 = dict(dtype=torch.uint8, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        color = torch.empty((3, H, W), **f32)
        normal = torch.empty((3, H, W), **f32)
        depth = torch.empty((1, H, W), **f32)
        opac = torch.empty((1, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=device)
        active = torch.empty((P,), dtype=torch.bool, device=device)

        st = ForwardState()
        st.frame, st.keep, st.radii, st.tile_mask, st.saved = frame, keep, radii, tile_mask, bool(save)
        gb, ib, _ = workspace_sizes(P, W, H, 0)
        st.geom = torch.empty((gb,), **u8)
        st.img = torch.empty((ib,), **u8)

        capacity = config.capacity if capacity is None else capacity
        auto_key = None
        if capacity == "auto":
            auto_key = (device.index, P, W, H)
            hist = _auto_history.get(auto_key)
            capacity = "exact" if not hist else int(max(hist) * config.auto_headroom) + config.auto_slack
        exact = capacity == "exact"
        if capturing and exact:
            raise RuntimeError(
                "eggsplat: a forward recorded into a CUDA graph needs a fixed binning capacity (no host read-back is "
                "possible inside a graph): set eggfusion_b200.rasterizer.config.capacity to an int, or to 'auto' and "
                "run one eager forward of this shape first")
        host = None if capturing else _get_pinned()
        own_first, own_count = (0, P) if own_range is None else (int(own_range[0]), int(own_range[1]))
        _lib.check(lib.egs_forward_plan_sharded(C.byref(frame), _ptr(means3D), _ptr(shs) if use_sh else None,
                                                None if use_sh else _ptr(colors_precomp), _ptr(opacities), _ptr(scales),
                                                _ptr(rotations), _ptr(tile_mask), own_first, own_count,
                                                st.geom.data_ptr(), st.img.data_ptr(), _ptr(radii), _ptr(active),
                                                host.data_ptr() if exact else None, stream), "forward_plan")
        if exact:
            torch.cuda.current_stream(device).synchronize()
            st.num_rendered, st.tile_num = int(host[0]), int(host[1])
            st.cap = st.num_rendered
            _pinned_free.append(host)
            if auto_key is not None:
                _auto_note(auto_key, st.num_rendered)
        else:
            st.cap = int(capacity)
            st.num_rendered = st.tile_num = -1  # unknown on the host by design
        bb = workspace_sizes(P, W, H, st.cap)[2] if save else _bin_bytes_forward_only(st.cap)
        st.bin = torch.empty((bb,), **u8)
        _lib.check(lib.egs_forward_render(C.byref(frame), _ptr(tile_mask), _ptr(radii), st.geom.data_ptr(),
                                          st.img.data_ptr(), st.bin.data_ptr(), st.cap, color.data_ptr(),
                                          normal.data_ptr(), depth.data_ptr(), opac.data_ptr(),
                                          None if (exact or capturing) else host.data_ptr(),
                                          0 if save else _lib.EGS_FWD_NO_SAVE, stream), "forward_render")
        if capturing:
            _captured.append((st.img[:16].view(torch.int32), st.cap))
        elif not exact:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            _pending_overflow.append((ev, host, st.cap, auto_key))
    return color, normal, depth, opac, active, radii, st


def backward_raw(st: ForwardState, means3D, shs, colors_precomp, scales, rotations, g_color, g_normal, g_depth,
                 g_opac, surfel_range=None, want_aux=False):
    """Run the CUDA backward.  Returns a dict of gradients (full [P,...] tensors)."""
    lib = _lib.load()
    device = means3D.device
    P = means3D.size(0)
    frame = st.frame
    if not getattr(st, "saved", True):
        raise RuntimeError("eggsplat: this forward was a forward-only render (save=False / no_grad): no backward state")
    with torch.cuda.device(device):
        if st.num_rendered == -1 and not _capturing():
            # the forward ran without a host read-back: if its counters have arrived by now (they usually have: the
            # loss was enqueued in between) an overflow is raised HERE, before any gradient of the truncated lists
            # reaches an optimiser; otherwise at the next forward (one step late, documented in _Config)
            _check_pending()
        stream = _stream_ptr(device)
        f32 = dict(dtype=torch.float32, device=device)
        means3D = _f32c(means3D, device)
        use_sh = _present(shs)
        M = frame.sh_coeffs
        shs = _f32c(shs, device) if use_sh else None
        colors_precomp = None if use_sh else _f32c(colors_precomp, device)
        scales, rotations = _f32c(scales, device), _f32c(rotations, device)
        g_color, g_normal = _f32c(g_color, device), _f32c(g_normal, device)
        g_depth, g_opac = _f32c(g_depth, device), _f32c(g_opac, device)
        out = {
            "means3D": torch.empty((P, 3), **f32), "opacities": torch.empty((P, 1), **f32),
            "sh": torch.empty((P, M, 3), **f32) if use_sh else None, "scales": torch.empty((P, 3), **f32),
            "rotations": torch.empty((P, 4), **f32),
            "colors_precomp": None if use_sh else torch.empty((P, 3), **f32),
            "means2D": torch.empty((P, 3), **f32) if want_aux else None,
            "cov3D": torch.empty((P, 6), **f32) if want_aux else None,
            "colors": torch.empty((P, 3), **f32) if (want_aux and use_sh) else None,
        }
        if P == 0:
            return out
        sg = torch.empty((P, _lib.SCREEN_GRAD_STRIDE), **f32)
        out["screen"] = sg
        _lib.check(lib.egs_backward_render(C.byref(frame), st.geom.data_ptr(), st.img.data_ptr(), st.bin.data_ptr(),
                                           st.cap, g_color.data_ptr(), g_normal.data_ptr(), g_depth.data_ptr(),
                                           g_opac.data_ptr(), sg.data_ptr(), 0, stream), "backward_render")
        first, count = (0, P) if surfel_range is None else surfel_range
        d_colors = out["colors_precomp"] if not use_sh else out["colors"]
        _lib.check(lib.egs_backward_surfels(C.byref(frame), first, count, means3D.data_ptr(), _ptr(shs),
                                            _ptr(colors_precomp), scales.data_ptr(), rotations.data_ptr(),
                                            st.radii.data_ptr(), st.geom.data_ptr(), sg.data_ptr(),
                                            out["means3D"].data_ptr(), out["opacities"].data_ptr(), _ptr(out["sh"]),
                                            out["scales"].data_ptr(), out["rotations"].data_ptr(),
                                            _ptr(out["means2D"]), _ptr(d_colors), _ptr(out["cov3D"]), stream),
                   "backward_surfels")
    return out


# ------------------------------------------------------------------------------------------------- autograd
class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, tile_mask,
                raster_settings):
        if _present(cov3Ds_precomp):
            # The reference dereferences rotations[idx] unconditionally (forward.cu:216) while its python front
            # end forces rotations to be absent whenever cov3D_precomp is given: that path cannot execute there.
            raise NotImplementedError("cov3D_precomp is not executable in the reference rasterizer (forward.cu:216)")
        args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, tile_mask, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, raster_settings.cx, raster_settings.cy, sh,
                raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                color, normal, depth, opac, active_mask, radii, st = forward_raw(
                    raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, tile_mask)
                torch.cuda.synchronize(means3D.device)  # debug mode surfaces kernel errors here (auxiliary.h:292-299)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            color, normal, depth, opac, active_mask, radii, st = forward_raw(
                raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, tile_mask)
        ctx.raster_settings = raster_settings
        ctx.state = st
        ctx.num_rendered = st.num_rendered
        ctx.num_tile = st.tile_num
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, sh)
        ctx.mark_non_differentiable(active_mask, radii)
        return color, normal, depth, opac, active_mask, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_normal, grad_out_depth, grad_out_opac, grad_active_mask, _):
        colors_precomp, means3D, scales, rotations, sh = ctx.saved_tensors
        st = ctx.state
        settings = ctx.raster_settings
        if settings.debug:
            cpu_args = cpu_deep_copy_tuple((means3D, colors_precomp, scales, rotations, sh, grad_out_color,
                                            grad_out_normal, grad_out_depth, grad_out_opac))
            try:
                g = backward_raw(st, means3D, sh, colors_precomp, scales, rotations, grad_out_color, grad_out_normal,
                                 grad_out_depth, grad_out_opac)
                torch.cuda.synchronize(means3D.device)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            g = backward_raw(st, means3D, sh, colors_precomp, scales, rotations, grad_out_color, grad_out_normal,
                             grad_out_depth, grad_out_opac)
        # same slots as the reference (__init__.py:152-162); absent inputs get None
        return (g["means3D"], g["sh"], g["colors_precomp"], g["opacities"], g["scales"], g["rotations"], None, None,
                None)


def rasterize_gaussians(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, tile_mask,
                        raster_settings):
    needs_grad = torch.is_grad_enabled() and any(
        isinstance(t, torch.Tensor) and t.requires_grad for t in (means3D, sh, colors_precomp, opacities, scales, rotations))
    if not needs_grad and not raster_settings.debug and not _present(cov3Ds_precomp):
        # the reference under torch.no_grad() (model maps for tracking / fusion, mapper.py:227,497): forward-only render
        color, normal, depth, opac, active_mask, radii, _st = forward_raw(
            raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, tile_mask, save=False)
        return color, normal, depth, opac, active_mask, radii
    return _RasterizeGaussians.apply(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     tile_mask, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean frustum mask (reference __init__.py:187-196)."""
        with torch.no_grad():
            s = self.raster_settings
            pos = _f32c(positions, positions.device)
            if not pos.is_cuda:
                raise RuntimeError("eggsplat: positions must be a CUDA tensor")
            P = pos.size(0)
            present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
            with torch.cuda.device(pos.device):
                view, proj = _f32c(s.viewmatrix, pos.device), _f32c(s.projmatrix, pos.device)
                _lib.check(_lib.load().egs_mark_visible(P, _ptr(pos), view.data_ptr(), proj.data_ptr(),
                                                        _ptr(present), _stream_ptr(pos.device)), "mark_visible")
        return present

    def forward(self, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, tile_mask=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        # tile_mask=None (the signature's default, which the reference binding cannot accept) means all tiles
        return rasterize_gaussians(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   tile_mask, raster_settings)


# ------------------------------------------------------------------------------------------------- test hook
def debug_export(st: ForwardState, P: int, W: int, H: int):
    """Copy the internal index artefacts out of a ForwardState (parity tests only)."""
    lib = _lib.load()
    dev = st.geom.device
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    o = {
        "point_list": torch.zeros((max(st.cap, 1),), **i32), "ranges": torch.zeros((tiles, 2), **i32),
        "tile_indices": torch.zeros((tiles,), **i32), "tiles_touched": torch.zeros((P,), **i32),
        "n_contrib": torch.zeros((H * W,), **i32), "final_T": torch.zeros((H * W,), **f32),
        "final_D": torch.zeros((H * W,), **f32), "records": torch.zeros((P, 16), **f32),
        "cov3D": torch.zeros((P, 6), **f32), "clamped": torch.zeros((P,), dtype=torch.uint8, device=dev),
    }
    with torch.cuda.device(dev):
        _lib.check(lib.egs_debug_export(C.byref(st.frame), st.geom.data_ptr(), st.img.data_ptr(), st.bin.data_ptr(),
                                        st.cap, o["point_list"].data_ptr(), o["ranges"].data_ptr(),
                                        o["tile_indices"].data_ptr(), _ptr(o["tiles_touched"]),
                                        o["n_contrib"].data_ptr(), o["final_T"].data_ptr(), o["final_D"].data_ptr(),
                                        _ptr(o["records"]), _ptr(o["cov3D"]), _ptr(o["clamped"]), _stream_ptr(dev)),
                   "debug_export")
    o["point_list"] = o["point_list"][:st.cap]
    return o
