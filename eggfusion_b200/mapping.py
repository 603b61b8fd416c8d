"""Mapping-iteration glue around the rasterizer (C ABI: include/eggmap.h).

Host-side mirror of the reference's per-iteration map optimisation
(/root/reference/src/core/mapper.py:336-368 `Mapper.frame_batch_optimization`):

    optimizer = torch.optim.Adam(self.surfels0.parametrize(self.sw_lr_params), lr=0.0)
    geo_surfels_params = {"position": get_xyz.detach(), "normal": get_normal.detach()}
    for it in range(iters):
        render_output = self.renderer.render(frame, self.total_params)       # activations + rasterizer forward
        loss = self.compute_loss(render_output, frame_map, masks, geo_surfels_params)
        loss.backward(); optimizer.step(); optimizer.zero_grad(set_to_none=True)

Two levels, same numbers:

  * `compute_loss` (autograd Function, image terms of Mapper.compute_loss) and `FrameBatchOptimizer`
    (`total_params`, `step`) slot into that loop as it stands -- the rasterizer stays `GaussianRasterizer`;
  * `FusedMapper.iterate` runs the whole iteration on persistent buffers with no autograd graph, no allocation and
    no host sync: plan/render -> egm_loss_seed -> backward_render -> backward_surfels -> egm_adam_step.

There is no CPU fallback: host tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch

from . import _lib
from . import rasterizer as R
from .pipeline import SplatContext


class MappingWeights(NamedTuple):
    """cfg.Mapping.{color,depth,normal,reg}_weight, reg_weight_n (mapper.py:150-154)."""
    color_weight: float = 1.0
    depth_weight: float = 1.0
    normal_weight: float = 1.0
    reg_weight: float = 0.0
    reg_weight_n: float = 1.0


class LrParams(NamedTuple):
    """The sw_lr_params / global_lr_params dicts of the reference (mapper.py:164-178)."""
    position_lr: float
    feature_lr: float
    opacity_lr: float
    scaling_lr: float
    rotation_lr: float


def _cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the mapping glue has no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32")
    return t.contiguous()


def _mask_u8(m: Optional[torch.Tensor], what: str):
    if m is None:
        return None
    if not m.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor")
    m = m.squeeze().contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)


def loss_seed(est_color, est_depth, est_normal, ref_color, ref_depth, ref_normal, rgb_mask, geo_mask,
              weights: MappingWeights, out=None, tile_mask=None):
    """egm_loss_seed on caller tensors.  Returns (terms[8] float64, dL_dcolor, dL_ddepth, dL_dnormal).
    `tile_mask` (int32 [tiles_y, tiles_x], a rank's share of a tile-sharded frame): sums and seeds only over the pixels
    of those tiles, the means' denominator stays the whole frame's (egm_loss_seed_tiles)."""
    lib = _lib.load()
    est_color, est_depth, est_normal = (_cuda_f32(est_color, "est_color"), _cuda_f32(est_depth, "est_depth"),
                                        _cuda_f32(est_normal, "est_normal"))
    H, W = est_color.shape[-2], est_color.shape[-1]
    ref_color = _cuda_f32(ref_color, "ref_color")
    ref_depth = None if ref_depth is None else _cuda_f32(ref_depth, "ref_depth")
    ref_normal = None if ref_normal is None else _cuda_f32(ref_normal, "ref_normal")
    rm, gm = _mask_u8(rgb_mask, "rgb_mask"), _mask_u8(geo_mask, "geo_mask")
    if ref_color.numel() != 3 * H * W or rm.numel() != H * W:
        raise RuntimeError("loss_seed: reference maps must be [H, W, C] of the rendered size")
    dev = est_color.device
    if out is None:
        out = (torch.empty(_lib.EGM_TERMS, dtype=torch.float64, device=dev), torch.empty_like(est_color),
               torch.empty_like(est_depth), torch.empty_like(est_normal))
    terms, gc, gd, gn = out
    with torch.cuda.device(dev):
        _lib.check(lib.egm_loss_seed_tiles(H, W, est_color.data_ptr(), est_depth.data_ptr(), est_normal.data_ptr(),
                                           ref_color.data_ptr(), R._ptr(ref_depth), R._ptr(ref_normal), rm.data_ptr(),
                                           None if gm is None else gm.data_ptr(), R._ptr(tile_mask),
                                           weights.color_weight, weights.depth_weight, weights.normal_weight,
                                           gc.data_ptr(), gd.data_ptr(), gn.data_ptr(), terms.data_ptr(),
                                           R._stream_ptr(dev)), "loss_seed")
    return terms, gc, gd, gn


def loss_total(terms, weights: MappingWeights, have_depth=True, have_normal=True, reg=None, step=0, P=0, out=None):
    """egm_loss_total: float32[5] = total, color, depth, normal, reg (device tensor; no sync)."""
    lib = _lib.load()
    dev = terms.device
    if out is None:
        out = torch.empty(5, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.egm_loss_total(terms.data_ptr(), None if reg is None else reg.data_ptr(), step, P,
                                      weights.color_weight, weights.depth_weight, weights.normal_weight,
                                      weights.reg_weight if reg is not None else 0.0, weights.reg_weight_n,
                                      int(have_depth), int(have_normal), out.data_ptr(), R._stream_ptr(dev)),
                   "loss_total")
    return out


class _MappingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, depth, normal, ref_color, ref_depth, ref_normal, rgb_mask, geo_mask, weights):
        terms, gc, gd, gn = loss_seed(color, depth, normal, ref_color, ref_depth, ref_normal, rgb_mask, geo_mask, weights)
        ctx.save_for_backward(gc, gd, gn)
        out = loss_total(terms, weights, ref_depth is not None, ref_normal is not None)
        ctx.mark_non_differentiable(terms)
        return out[0], terms

    @staticmethod
    def backward(ctx, g, _g_terms):
        gc, gd, gn = ctx.saved_tensors
        return gc * g, gd * g, gn * g, None, None, None, None, None, None


def compute_loss(render_output, frame_input, render_mask, weights: MappingWeights = MappingWeights()):
    """Image terms of Mapper.compute_loss (mapper.py:381-426,437-438) as ONE kernel pair, autograd-connected to the
    rendered color / depth / normal.  `render_output`: dict color [3,H,W], depth [1,H,W], normal [3,H,W];
    `frame_input`: dict color_map [H,W,3], depth_map [H,W,1] | None, normal_map_c [H,W,3] | None;
    `render_mask`: (rgb_mask, geo_mask).  The regulariser term lives in FrameBatchOptimizer.step."""
    rgb_mask, geo_mask = render_mask
    loss, _terms = _MappingLoss.apply(render_output["color"], render_output["depth"], render_output["normal"],
                                      frame_input["color_map"], frame_input.get("depth_map"),
                                      frame_input.get("normal_map_c"), rgb_mask, geo_mask, weights)
    return loss


class FrameBatchOptimizer:
    """Owns the raw surfel parameters, their activations and the Adam state of one optimisation window.

    `surfels`: an object with the GaussianSurfels tensors `_xyz [P,3]`, `_features_dc [P,1,3]`,
    `_features_rest [P,M-1,3]`, `_scaling [P,3]`, `_rotation [P,4]`, `_opacity [P,1]` (gaussian_surfels.py:16-32), or a
    dict with the same keys without the underscore.  The values are copied in; `write_back()` copies them out."""

    NAMES = ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")

    def __init__(self, surfels, lr: LrParams, weights: MappingWeights = MappingWeights(), betas=(0.9, 0.999),
                 eps: float = 1e-8, padded_rows: Optional[int] = None):
        """padded_rows (>= P): allocate the five tensors the renderer consumes with that many rows (the attributes stay
        [P, ...] views), so that a sharded optimiser can all-gather them in place (parallel.DistributedMapper)."""
        self.lib = _lib.load()
        self.surfels = surfels
        get = (lambda n: surfels[n]) if isinstance(surfels, dict) else (lambda n: getattr(surfels, "_" + n))
        src = {n: get(n).detach() for n in self.NAMES}
        for n, t in src.items():
            _cuda_f32(t, n)
        dev = src["xyz"].device
        self.device = dev
        self.P = int(src["xyz"].shape[0])
        self.M = 1 + int(src["features_rest"].shape[1])
        self.lr, self.weights, self.betas, self.eps = lr, weights, betas, eps
        P, M = self.P, self.M
        f32 = dict(dtype=torch.float32, device=dev)
        rows = P if padded_rows is None else max(int(padded_rows), P)
        self._full = {}

        def alloc(name, *shape):
            self._full[name] = torch.zeros((rows,) + shape, **f32)
            return self._full[name][:P]
        # raw parameters; SH rows 0 / 1.. are _features_dc / _features_rest (identity activation: raw == activated)
        self.xyz = alloc("xyz", 3)
        self.xyz.copy_(src["xyz"])
        self.shs = alloc("shs", M, 3)
        self.shs[:, :1].copy_(src["features_dc"])
        self.shs[:, 1:].copy_(src["features_rest"])
        self.scaling_raw = src["scaling"].clone().contiguous()
        self.rotation_raw = src["rotation"].clone().contiguous()
        self.opacity_raw = src["opacity"].clone().contiguous()
        # activated
        self.opacity, self.scales = alloc("opacity", 1), alloc("scales", 3)
        self.rotations, self.normal0 = alloc("rotations", 4), torch.empty((P, 3), **f32)
        # Adam state
        self.state = {n: (torch.zeros_like(t), torch.zeros_like(t)) for n, t in
                      (("xyz", self.xyz), ("shs", self.shs), ("opacity", self.opacity_raw),
                       ("scaling", self.scaling_raw), ("rotation", self.rotation_raw))}
        self.step_count = 0
        self.reg = torch.zeros(_lib.EGM_REG, dtype=torch.float64, device=dev)
        self.loss_out = torch.zeros(5, **f32)
        with torch.cuda.device(dev):
            _lib.check(self.lib.egm_activate(P, self.opacity_raw.data_ptr(), self.scaling_raw.data_ptr(),
                                             self.rotation_raw.data_ptr(), self.opacity.data_ptr(),
                                             self.scales.data_ptr(), self.rotations.data_ptr(), self.normal0.data_ptr(),
                                             R._stream_ptr(dev)), "activate")
        self.pos0 = self.xyz.clone()          # geo_surfels_params["position"], mapper.py:342-345
        self._leaves = None

    # ---- Mapper.total_params (mapper.py:565-585): what Renderer.render consumes
    @property
    def total_params(self):
        """Activated parameters as autograd leaves (their .grad is what `step()` consumes)."""
        if self._leaves is None:
            self._leaves = {"xyz": self.xyz.requires_grad_(True), "opacity": self.opacity.requires_grad_(True),
                            "scales": self.scales.requires_grad_(True), "rotations": self.rotations.requires_grad_(True),
                            "shs": self.shs.requires_grad_(True)}
        return self._leaves

    def _hyper(self):
        lr, w = self.lr, self.weights
        return _lib.AdamHyper(self.betas[0], self.betas[1], self.eps, lr.position_lr, lr.feature_lr,
                              lr.feature_lr / 20.0, lr.opacity_lr, lr.scaling_lr, lr.rotation_lr, self.step_count,
                              w.reg_weight, w.reg_weight_n)

    def step(self, grads=None, first: int = 0, count: Optional[int] = None, _sh_done: bool = False) -> None:
        """optimizer.step() + optimizer.zero_grad(): `grads` = dict xyz, shs, opacity, scales, rotations of gradients
        w.r.t. the activated parameters (default: the .grad of `total_params`).
        `first`, `count`: update only the surfel rows [first, first + count) (a rank's owned range in a sharded
        optimisation, parallel.DistributedMapper); the regulariser's partial sums in `self.reg` then cover that range
        only and the caller adds the ranks' parts."""
        if grads is None:
            tp = self.total_params
            zero = lambda t: torch.zeros_like(t)
            grads = {k: (tp[k].grad if tp[k].grad is not None else zero(tp[k])) for k in tp}
        g = {k: _cuda_f32(v, "grad " + k) for k, v in grads.items()}
        if not _sh_done:
            self.step_count += 1
        h = self._hyper()
        st = self.state
        dev = self.device
        count = self.P - first if count is None else int(count)
        if count < self.P and count > 0:
            # the kernel averages the normal regulariser over the rows it is given: keep the mean over all P surfels
            h.reg_weight_n = h.reg_weight_n * (count / self.P)
        M3 = self.M * 3
        at = lambda t, width: t.data_ptr() + 4 * width * first      # row `first` of a [P, width] fp32 array
        sh = not _sh_done
        with torch.cuda.device(dev), torch.no_grad():
            _lib.check(self.lib.egm_adam_step(
                count, self.M if sh else 0, C.byref(h), at(self.xyz, 3), at(self.shs, M3) if sh else None,
                at(self.opacity_raw, 1),
                at(self.scaling_raw, 3), at(self.rotation_raw, 4), at(g["xyz"], 3), at(g["shs"], M3) if sh else None,
                at(g["opacity"], 1), at(g["scales"], 3), at(g["rotations"], 4),
                at(st["xyz"][0], 3), at(st["xyz"][1], 3), at(st["shs"][0], M3) if sh else None,
                at(st["shs"][1], M3) if sh else None,
                at(st["opacity"][0], 1), at(st["opacity"][1], 1), at(st["scaling"][0], 3),
                at(st["scaling"][1], 3), at(st["rotation"][0], 4), at(st["rotation"][1], 4),
                at(self.pos0, 3), at(self.normal0, 3), self.reg.data_ptr(), at(self.opacity, 1),
                at(self.scales, 3), at(self.rotations, 4), R._stream_ptr(dev)), "adam_step")
        if self._leaves is not None:
            for t in self._leaves.values():
                t.grad = None

    @property
    def can_fuse_sh(self) -> bool:
        """egm_backward_surfels_adam covers the layout the mapping loop uses: 16 SH coefficients, 16-byte aligned."""
        st = self.state["shs"]
        return self.M == 16 and all(t.data_ptr() % 16 == 0 for t in (self.shs, st[0], st[1]))

    def backward_surfels_and_step(self, ctx, first: int = 0, count: Optional[int] = None,
                                  screen_base: Optional[int] = None, mark=None) -> None:
        """The per-surfel backward of `ctx`'s frame and optimizer.step() as two launches: k_surfel_backward applies
        Adam to the SH block while the gradient rows are still in shared memory (egm_backward_surfels_adam), the other
        four groups follow in k_adam_geom.  Same results as ctx.backward_surfels() + step() (bit-identical: the same
        update function on the same gradient)."""
        count = self.P - first if count is None else int(count)
        self.step_count += 1
        h = self._hyper()
        st = self.state["shs"]
        with torch.cuda.device(self.device), torch.no_grad():
            _lib.check(self.lib.egm_backward_surfels_adam(
                C.byref(ctx.frame), first, count, self.xyz.data_ptr(), self.shs.data_ptr(), self.scales.data_ptr(),
                self.rotations.data_ptr(), ctx.radii.data_ptr(), ctx.geom.data_ptr(),
                ctx.screen.data_ptr() if screen_base is None else screen_base, ctx.d_means.data_ptr(),
                ctx.d_opac.data_ptr(), ctx.d_scales.data_ptr(), ctx.d_rots.data_ptr(), C.byref(h), st[0].data_ptr(),
                st[1].data_ptr(), R._stream_ptr(self.device)), "backward_surfels_adam")
        if mark:
            mark("bwd_surfels")
        self.step({"xyz": ctx.d_means, "opacity": ctx.d_opac, "scales": ctx.d_scales, "rotations": ctx.d_rots},
                  first=first, count=count, _sh_done=True)

    def loss_values(self, terms, have_depth=True, have_normal=True):
        """float32[5] device tensor: total, color, depth, normal, reg of the iteration just stepped (no host sync)."""
        return loss_total(terms, self.weights, have_depth, have_normal, self.reg if self.weights.reg_weight > 0 else None,
                          self.step_count, self.P, self.loss_out)

    def raw_params(self):
        return {"xyz": self.xyz, "features_dc": self.shs[:, :1], "features_rest": self.shs[:, 1:],
                "scaling": self.scaling_raw, "rotation": self.rotation_raw, "opacity": self.opacity_raw}

    def write_back(self) -> None:
        """Copy the optimised raw parameters back into the `surfels` object given at construction."""
        raw = self.raw_params()
        with torch.no_grad():
            for n in self.NAMES:
                dst = self.surfels[n] if isinstance(self.surfels, dict) else getattr(self.surfels, "_" + n)
                dst.data.copy_(raw[n])


class FusedMapper:
    """One mapping iteration = 4 C-ABI stages on persistent buffers (SplatContext) + 2 glue stages, no autograd, no
    allocation, no host sync.  `iterate()` returns the device tensor [total, color, depth, normal, reg].
    Capacity overflow (more (tile, surfel) instances than the fixed binning workspace holds) is watched without a sync:
    every frame's counters follow the render to pinned host memory and the NEXT iterate() (or synchronize()) raises if
    one of them carried the overflow flag."""

    def __init__(self, opt: FrameBatchOptimizer, width: int, height: int, capacity: int, sh_degree: int,
                 fuse_sh_adam: Optional[bool] = None):
        """fuse_sh_adam: apply Adam to the SH block inside the per-surfel backward kernel (default: whenever the layout
        allows it, FrameBatchOptimizer.can_fuse_sh); False keeps the two separate passes."""
        self.opt = opt
        self.fuse_sh_adam = opt.can_fuse_sh if fuse_sh_adam is None else bool(fuse_sh_adam)
        self.ctx = SplatContext(opt.P, width, height, opt.M, capacity, device=opt.device)
        self.sh_degree = sh_degree
        dev = opt.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.terms = torch.zeros(_lib.EGM_TERMS, dtype=torch.float64, device=dev)
        self.g_color, self.g_normal = torch.empty((3, height, width), **f32), torch.empty((3, height, width), **f32)
        self.g_depth, self.g_opac = torch.empty((1, height, width), **f32), torch.zeros((1, height, width), **f32)

    def iterate(self, settings, frame_input, render_mask, mark=None):
        o, ctx = self.opt, self.ctx
        ctx.check_overflow()
        ctx.set_camera(settings)
        with torch.no_grad():
            ctx.forward(o.xyz, o.shs, None, o.opacity, o.scales, o.rotations, None, mark, watch_overflow=True)
            loss_seed(ctx.color, ctx.depth, ctx.normal, frame_input["color_map"], frame_input.get("depth_map"),
                      frame_input.get("normal_map_c"), render_mask[0], render_mask[1], o.weights,
                      out=(self.terms, self.g_color, self.g_depth, self.g_normal))
            if mark:
                mark("loss_seed")
            ctx.backward_render(self.g_color, self.g_normal, self.g_depth, self.g_opac, mark=mark)
            if self.fuse_sh_adam:
                o.backward_surfels_and_step(ctx, mark=mark)
            else:
                ctx.backward_surfels(o.xyz, o.shs, None, o.scales, o.rotations, mark=mark)
                o.step({"xyz": ctx.d_means, "shs": ctx.d_sh, "opacity": ctx.d_opac, "scales": ctx.d_scales,
                        "rotations": ctx.d_rots})
            if mark:
                mark("adam")
            return o.loss_values(self.terms, frame_input.get("depth_map") is not None,
                                 frame_input.get("normal_map_c") is not None)

    def synchronize(self) -> None:
        """Waits for the iterations issued so far and raises if any of them overflowed the binning capacity."""
        torch.cuda.synchronize(self.opt.device)
        self.ctx.check_overflow(block=True)
