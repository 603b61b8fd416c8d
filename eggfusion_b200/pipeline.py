"""Persistent, allocation-free and host-sync-free forward+backward context over the C-ABI.

`GaussianRasterizer` (rasterizer.py) mirrors the reference API and therefore allocates fresh outputs per call and
reads the instance count back once per forward.  A mapping loop that renders the same surfel set from a few cameras
hundreds of times (/root/reference/src/core/mapper.py:336-368) can instead keep one `SplatContext`: all workspaces
and outputs are allocated once for (P, W, H, capacity) and every step is a fixed sequence of stream-ordered
launches with no host round trip -- the shape CUDA graphs want.  bench.py's device-resident leg uses it and also
hooks per-stage CUDA events into it.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib
from . import rasterizer as R


class SplatContext:
    STAGES = ("plan", "render", "bwd_zero", "bwd_render", "bwd_surfels")

    def __init__(self, P: int, width: int, height: int, sh_coeffs: int, capacity: int, device="cuda:0",
                 padded_rows: Optional[int] = None, own_range=None):
        """own_range = (first, count): this context is one rank of a tile-sharded frame that owns those surfel rows
        (egs_forward_plan_sharded skips the colour / record work for surfels nobody on this rank reads)."""
        self.lib = _lib.load()
        self.P, self.W, self.H, self.M, self.cap = int(P), int(width), int(height), int(sh_coeffs), int(capacity)
        self.device = torch.device(device)
        dev = self.device
        with torch.cuda.device(dev):
            gb, ib, bb = R.workspace_sizes(self.P, self.W, self.H, self.cap)
            u8 = dict(dtype=torch.uint8, device=dev)
            f32 = dict(dtype=torch.float32, device=dev)
            self.geom, self.img, self.bin = torch.empty(gb, **u8), torch.empty(ib, **u8), torch.empty(bb, **u8)
            self.color, self.normal = torch.empty((3, height, width), **f32), torch.empty((3, height, width), **f32)
            self.depth, self.opacity = torch.empty((1, height, width), **f32), torch.empty((1, height, width), **f32)
            self.radii = torch.empty((P,), dtype=torch.int32, device=dev)
            self.active = torch.empty((P,), dtype=torch.bool, device=dev)
            rows = self.P if padded_rows is None else int(padded_rows)
            self.screen = torch.zeros((rows, _lib.SCREEN_GRAD_STRIDE), **f32)
            self.d_means = torch.empty((P, 3), **f32)
            self.d_opac = torch.empty((P, 1), **f32)
            self.d_sh = torch.empty((P, max(sh_coeffs, 1), 3), **f32)
            self.d_scales = torch.empty((P, 3), **f32)
            self.d_rots = torch.empty((P, 4), **f32)
            self.d_colors = torch.empty((P, 3), **f32)
            self.counters_host = torch.zeros((4,), dtype=torch.int32).pin_memory()
        self.frame = None
        self._keep = None
        self._watch = []        # (event, pinned counters) of forwards whose overflow flag has not been looked at yet
        self.own_range = (0, self.P) if own_range is None else (int(own_range[0]), int(own_range[1]))

    def set_camera(self, settings) -> None:
        self.frame, self._keep = R.make_frame(self.P, settings, self.M, self.device)

    # each stage is one C-ABI call; `mark(stage_name)` (optional) is invoked after each for event timing
    def forward(self, means3D, shs, colors_precomp, opacities, scales, rotations, tile_mask=None,
                mark: Optional[Callable[[str], None]] = None, fetch_counters: bool = False, save: bool = True,
                watch_overflow: bool = False) -> None:
        """watch_overflow: the frame's counters travel to a pinned host buffer behind the render (asynchronously, no
        sync) and `check_overflow()` raises once they have arrived with the overflow flag set -- the fixed-capacity
        binning workspace truncated the instance lists, so images and gradients of that frame are wrong.  Not available
        while the stream is being captured into a CUDA graph (read_counters() after the replay instead)."""
        lib, fr = self.lib, self.frame
        stream = R._stream_ptr(self.device)
        tm = None if tile_mask is None else tile_mask.data_ptr()
        watch = None
        if watch_overflow and not R._capturing():
            watch = R._get_pinned()
        with torch.cuda.device(self.device):
            _lib.check(lib.egs_forward_plan_sharded(C.byref(fr), means3D.data_ptr(), R._ptr(shs), R._ptr(colors_precomp),
                                                    opacities.data_ptr(), scales.data_ptr(), rotations.data_ptr(), tm,
                                                    self.own_range[0], self.own_range[1], self.geom.data_ptr(),
                                                    self.img.data_ptr(), self.radii.data_ptr(), self.active.data_ptr(),
                                                    None, stream), "forward_plan")
        if mark:
            mark("plan")
        with torch.cuda.device(self.device):
            _lib.check(lib.egs_forward_render(C.byref(fr), tm, self.radii.data_ptr(), self.geom.data_ptr(),
                                              self.img.data_ptr(), self.bin.data_ptr(), self.cap, self.color.data_ptr(),
                                              self.normal.data_ptr(), self.depth.data_ptr(), self.opacity.data_ptr(),
                                              watch.data_ptr() if watch is not None else
                                              (self.counters_host.data_ptr() if fetch_counters else None),
                                              0 if save else _lib.EGS_FWD_NO_SAVE, stream),
                       "forward_render")
            if watch is not None:
                ev = torch.cuda.Event()
                ev.record()
                self._watch.append((ev, watch))
        if mark:
            mark("render")

    def check_overflow(self, block: bool = False) -> None:
        """Raises if a watched forward (forward(watch_overflow=True)) overflowed the binning capacity.  Non-blocking by
        default: only counters that have already arrived are examined; block=True waits for all of them."""
        while self._watch:
            ev, host = self._watch[0]
            if not block and not ev.query():
                return
            ev.synchronize()
            self._watch.pop(0)
            n, flag = int(host[0]), int(host[2])
            R._pinned_free.append(host)
            if flag != 0:
                self._watch.clear()
                raise RuntimeError(
                    f"eggsplat: a frame produced {n} (tile, surfel) instances but this SplatContext's binning workspace "
                    f"holds {self.cap}; its lists were truncated and that frame's images / gradients are wrong. "
                    f"Build the context (FusedMapper / DistributedMapper) with a larger capacity.")

    def backward_render(self, g_color, g_normal, g_depth, g_opac, mark=None, prezeroed: bool = False) -> None:
        """prezeroed: the screen-gradient block is known to be all zeros already (the peer exchange of a sharded step
        clears every row it pushes, parallel.PeerExchange), so the 64 B/surfel memset is skipped."""
        lib, fr = self.lib, self.frame
        stream = R._stream_ptr(self.device)
        with torch.cuda.device(self.device):
            if not prezeroed:
                self.screen[:self.P].zero_()
            if mark:
                mark("bwd_zero")
            _lib.check(lib.egs_backward_render(C.byref(fr), self.geom.data_ptr(), self.img.data_ptr(),
                                               self.bin.data_ptr(), self.cap, g_color.data_ptr(), g_normal.data_ptr(),
                                               g_depth.data_ptr(), g_opac.data_ptr(), self.screen.data_ptr(),
                                               _lib.EGS_BWD_GRADS_PREZEROED, stream), "backward_render")
        if mark:
            mark("bwd_render")

    def backward_surfels(self, means3D, shs, colors_precomp, scales, rotations, first=0, count=None,
                         screen_base: Optional[int] = None, mark=None) -> None:
        lib, fr = self.lib, self.frame
        stream = R._stream_ptr(self.device)
        count = self.P - first if count is None else count
        use_sh = shs is not None
        with torch.cuda.device(self.device):
            _lib.check(lib.egs_backward_surfels(C.byref(fr), first, count, means3D.data_ptr(), R._ptr(shs),
                                                R._ptr(colors_precomp), scales.data_ptr(), rotations.data_ptr(),
                                                self.radii.data_ptr(), self.geom.data_ptr(),
                                                self.screen.data_ptr() if screen_base is None else screen_base,
                                                self.d_means.data_ptr(), self.d_opac.data_ptr(),
                                                self.d_sh.data_ptr() if use_sh else None, self.d_scales.data_ptr(),
                                                self.d_rots.data_ptr(), None,
                                                None if use_sh else self.d_colors.data_ptr(), None, stream),
                       "backward_surfels")
        if mark:
            mark("bwd_surfels")

    def step(self, params: dict, pixel_grads, tile_mask=None, mark=None) -> None:
        """forward + full backward for params = {xyz, opacity, shs, scales, rotations} (contiguous fp32 CUDA)."""
        self.forward(params["xyz"], params.get("shs"), params.get("colors"), params["opacity"], params["scales"],
                     params["rotations"], tile_mask, mark)
        self.backward_render(*pixel_grads, mark=mark)
        self.backward_surfels(params["xyz"], params.get("shs"), params.get("colors"), params["scales"],
                              params["rotations"], mark=mark)

    def read_counters(self):
        """Blocking read of the device counters (num_rendered, tile_num, overflow, num_visible)."""
        c = self.img[:16].view(torch.int32).cpu()
        return tuple(int(v) for v in c)
