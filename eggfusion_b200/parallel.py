"""Multi-GPU sharding of one render/optimise step (one process per GPU, torch.distributed for the plumbing).

The reference is single-GPU; its `tile_mask` argument (forward.cu:292-300, rasterizer_impl.cu:103-111) is the seam:
a surfel only emits instances for tiles whose mask is non-zero.  Scheme:

  * every rank holds the full (activated) surfel parameter set.  The exact per-surfel projection, the colour (SH)
    evaluation and the record / backward state are produced only for surfels that can reach one of the rank's tiles or
    lie in its owned surfel range (`egs_forward_plan_sharded`: from 3 ranks on a cheap conservative footprint bound
    first compacts those candidates, the projection then runs on them alone);
  * rank r bins / sorts / composites only its tile set -- a contiguous run of the row-major tile sequence with 1/world
    of the summed cost (per-tile list lengths of previous frames; `tile_partition`), or interleaved tile rows when no
    costs are known -- and runs the reverse walk over the same tiles -> partial rows of the screen-gradient block
    G[P][16] for the ~P/world surfels it touched;
  * ONE exchange (`PeerExchange`): the rows a rank touched are compacted per 256-surfel group and stored into its
    section of their owners' inboxes in NVLink peer memory (`egs_push_rows`: one TMA bulk store per group on the
    peer-mapped address; symmetric memory from torch.distributed._symmetric_memory), a device-side barrier follows, and
    every owner folds its inbox sections into its block (`egs_fold_inbox`).  Only touched rows cross the links (C3 at 8
    GPUs: ~12 MB per rank instead of the 56 MB of a dense reduce-scatter).  Without peer memory (gloo, no P2P) the
    same sum is one `reduce_scatter_tensor` over NCCL;
  * rank r runs the per-surfel backward for its surfel range -> gradient shards [first, first + count);
  * a distributed optimiser step (`DistributedMapper`): Adam on the owned range, then `all_gather_into_tensor` of the
    updated activated parameters.

`ShardedRasterizer` is the autograd-level entry (the reference API's `GaussianRasterizer`, sharded): the exchange runs
inside `backward`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from ._lib import SCREEN_GRAD_STRIDE


# ------------------------------------------------------------------------------------------------- partitions
def tile_partition(tiles_y: int, tiles_x: int, world: int, rank: int,
                   costs: Optional[torch.Tensor] = None, layout: str = "bands") -> torch.Tensor:
    """int32 [tiles_y, tiles_x] mask of the tiles rank `rank` renders.  Masks of all ranks are disjoint and cover
    the grid.  Without `costs`: tile rows are dealt round-robin.  With `costs` (per-tile list lengths, any device):
      layout "bands" (default): the row-major tile sequence is cut into `world` CONTIGUOUS runs of equal summed cost
        (balance to one tile).  A surfel's tile rectangle then meets one rank, two at a band edge -- so the sharded
        projection keeps ~1/world of the surfels per rank (egs_forward_plan_sharded) and a surfel's screen-gradient
        row is pushed to its owner once, not once per tile row it spans;
      layout "rows": whole tile rows assigned greedily, longest first (rows of one rank are scattered over the image)."""
    mask = torch.zeros((tiles_y, tiles_x), dtype=torch.int32)
    if costs is None:
        mask[rank::world, :] = 1
        return mask
    c = costs.detach().to("cpu", torch.float64).reshape(tiles_y, tiles_x)
    if layout == "bands":
        # a tile costs its list length plus a constant (an empty tile still occupies a CTA slot for a moment)
        flat = c.reshape(-1) + 1.0
        cum = torch.cumsum(flat, 0)
        total = float(cum[-1])
        mid = cum - 0.5 * flat                                  # a tile belongs to the run its midpoint falls in
        owner = torch.clamp((mid * (world / total)).floor().long(), 0, world - 1)
        return (owner == rank).to(torch.int32).reshape(tiles_y, tiles_x)
    if layout != "rows":
        raise ValueError("tile_partition: layout must be 'bands' or 'rows'")
    row_cost = c.sum(dim=1)
    order = torch.argsort(row_cost, descending=True, stable=True).tolist()
    load = [0.0] * world
    owner = [0] * tiles_y
    for r in order:
        w = min(range(world), key=lambda k: (load[k], k))
        owner[r] = w
        load[w] += float(row_cost[r])
    for r in range(tiles_y):
        if owner[r] == rank:
            mask[r, :] = 1
    return mask


CHUNK_ALIGN = 256   # rows; a rank's surfel range starts at a multiple of it (egs_push_rows works on 256-surfel groups)


def padded_rows(P: int, world: int) -> int:
    """Rows of the exchanged block: every rank owns the same number of rows (`padded_rows // world`), a multiple of
    CHUNK_ALIGN when the block is split at all."""
    if world <= 1:
        return P
    chunk = (P + world - 1) // world
    chunk = (chunk + CHUNK_ALIGN - 1) // CHUNK_ALIGN * CHUNK_ALIGN
    return chunk * world


def surfel_range(P: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of the surfel rows rank `rank` owns after the exchange."""
    chunk = padded_rows(P, world) // world
    first = min(P, rank * chunk)
    return first, max(0, min(P, first + chunk) - first)


def reduce_scatter_rows(block: torch.Tensor, group=None) -> torch.Tensor:
    """Sum `block` ([padded_rows, C], identical shape on all ranks) over the group and return this rank's
    contiguous row chunk.  NCCL: a single reduce_scatter_tensor.  Gloo (CPU tests) has no reduce-scatter, so the
    same result is produced with all_reduce + slice."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = block.shape[0]
    assert rows % world == 0, "block must be padded with padded_rows()"
    chunk = rows // world
    if world == 1:
        return block[:chunk]
    if dist.get_backend(group) == "nccl":
        out = torch.empty((chunk,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
        dist.reduce_scatter_tensor(out, block.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    tmp = block.clone()
    dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
    return tmp[rank * chunk:(rank + 1) * chunk].contiguous()


def all_gather_rows(t: torch.Tensor, first: int, count: int, chunk: int, world: int, group=None, cache=None,
                    full: Optional[torch.Tensor] = None) -> None:
    """In place: every rank's owned rows [first, first + count) of `t` ([P, ...], same shape everywhere) are
    distributed to all ranks.  `full`: the [world * chunk, ...] allocation `t` is the head of (FrameBatchOptimizer with
    padded_rows): one in-place all_gather_into_tensor, no staging.  Otherwise equal padded chunks are staged; gloo (CPU
    tests) gathers a list."""
    P = t.shape[0]
    flat = t.view(P, -1)
    nccl = dist.get_backend(group) == "nccl"
    if full is not None and nccl and full.shape[0] == world * chunk and full.data_ptr() == t.data_ptr():
        f2 = full.view(world * chunk, -1)
        rank = dist.get_rank(group)
        dist.all_gather_into_tensor(f2, f2[rank * chunk:(rank + 1) * chunk], group=group)
        return
    key = (flat.shape[1], flat.dtype)
    cache = {} if cache is None else cache
    if key not in cache:
        cache[key] = (torch.zeros((chunk, flat.shape[1]), dtype=flat.dtype, device=flat.device),
                      torch.empty((world * chunk, flat.shape[1]), dtype=flat.dtype, device=flat.device))
    mine, staged = cache[key]
    mine[:count].copy_(flat[first:first + count])
    if nccl:
        dist.all_gather_into_tensor(staged, mine, group=group)
    else:
        dist.all_gather(list(staged.view(world, chunk, -1).unbind(0)), mine, group=group)
    flat.copy_(staged[:P])


def _world_rank(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


# ------------------------------------------------------------------------------------------------- the exchange
class PeerExchange:
    """The screen-gradient exchange of a tile-sharded step over NVLink peer (symmetric) memory.

    Every rank owns an INBOX per parity (symmetric memory, peers store into it): a header of `world` row counts and
    [world senders][chunk rows][16 floats].  A step: `push` (egs_push_rows: the rows this rank's reverse walk touched
    are compacted per 256-surfel group and streamed into the owners' inboxes with coalesced stores; the local rows are
    cleared), `barrier` (device-side, all ranks), `fold` (egs_fold_inbox: the owner adds what the senders left into
    its block).  An inbox is written again two steps later, after at least one more barrier that its owner joins only
    after the fold (stream order), so one barrier per step is enough.
    """

    HEAD = 256   # bytes reserved for the header in front of the inbox

    def __init__(self, P: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.lib = _lib.load()
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.P = int(P)
        self.chunk = max(padded_rows(P, self.world) // self.world, CHUNK_ALIGN)
        self.first = min(self.P, self.rank * self.chunk)
        self.device = torch.device(device)
        name = self.group.group_name
        try:
            symm.enable_symm_mem_for_group(name)
        except Exception:
            pass
        nfl = self.HEAD // 4 + self.world * self.chunk * SCREEN_GRAD_STRIDE
        with torch.cuda.device(self.device):
            self.buffers = [symm.empty((nfl,), dtype=torch.float32, device=self.device) for _ in range(2)]
            self.handles = [symm.rendezvous(b, name) for b in self.buffers]
            for b in self.buffers:
                b[:self.HEAD // 4].zero_()
            i64 = dict(dtype=torch.int64, device=self.device)
            # device arrays of the world's header / inbox addresses as this process maps them
            self.headers = [torch.tensor([int(p) for p in h.buffer_ptrs], **i64) for h in self.handles]
            self.inboxes = [torch.tensor([int(p) + self.HEAD for p in h.buffer_ptrs], **i64) for h in self.handles]
            self.sent = torch.zeros((self.world,), dtype=torch.int32, device=self.device)
            self.block = torch.empty((self.chunk, SCREEN_GRAD_STRIDE), dtype=torch.float32, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        self.parity = 0

    def exchange(self, geom: torch.Tensor, local_sg: torch.Tensor, stream: int, mark=None) -> int:
        """push + barrier + fold.  Returns the address at which row 0 of the FULL [P][16] block would lie if the owned
        block were a window into it (the per-surfel kernel indexes by global surfel id)."""
        from . import _lib
        b = self.parity
        self.parity ^= 1
        _lib.check(self.lib.egs_push_rows(self.P, self.chunk, self.world, self.rank, geom.data_ptr(), local_sg.data_ptr(),
                                          self.sent.data_ptr(), self.inboxes[b].data_ptr(), self.headers[b].data_ptr(),
                                          stream), "push_rows")
        if mark:
            mark("push_rows")
        self.handles[b].barrier(channel=0)
        if mark:
            mark("barrier")
        buf = self.buffers[b]
        _lib.check(self.lib.egs_fold_inbox(self.chunk, self.world, self.first, buf.data_ptr() + self.HEAD, buf.data_ptr(),
                                           self.block.data_ptr(), stream), "fold_inbox")
        if mark:
            mark("fold_inbox")
        return self.block.data_ptr() - self.first * SCREEN_GRAD_STRIDE * 4

    def consumed(self) -> None:
        """Kept for symmetry with the reduce-scatter path: egs_fold_inbox rewrites the whole block."""


def make_exchange(P: int, device, group=None) -> Optional[PeerExchange]:
    """PeerExchange when the group runs on NCCL with peer access, else None (callers fall back to reduce-scatter)."""
    world, _ = _world_rank(group)
    if world == 1 or not torch.cuda.is_available() or dist.get_backend(group) != "nccl":
        return None
    import os
    if os.environ.get("EGS_EXCHANGE", "peer") != "peer":
        return None
    try:
        return PeerExchange(P, device, group)
    except Exception as ex:     # no symmetric memory on this system: NCCL reduce-scatter instead
        if dist.get_rank(group) == 0:
            print("eggfusion_b200.parallel: peer exchange unavailable (%s); using NCCL reduce-scatter" % (ex,))
        return None


# ------------------------------------------------------------------------------------------------- raw sharded step
class ShardedSplat:
    """One rank's share of a tile-sharded forward + backward (fresh tensors per call, like rasterizer.forward_raw).
    CUDA only."""

    def __init__(self, group=None, costs: Optional[torch.Tensor] = None, layout: str = "bands"):
        self.group = group
        self.world, self.rank = _world_rank(group)
        self.costs = costs
        self.layout = layout
        self._mask_cache = {}
        self._exchange = {}

    def mask_for(self, tiles_y: int, tiles_x: int, device) -> torch.Tensor:
        key = (tiles_y, tiles_x, str(device))
        if key not in self._mask_cache:
            self._mask_cache[key] = tile_partition(tiles_y, tiles_x, self.world, self.rank, self.costs, self.layout).to(device)
        return self._mask_cache[key]

    def set_costs(self, costs: Optional[torch.Tensor]) -> None:
        """New per-tile costs (e.g. the previous frame's list lengths): the tile masks are rebuilt on next use."""
        self.costs = costs
        self._mask_cache.clear()

    def exchange_for(self, P: int, device) -> Optional[PeerExchange]:
        key = (int(P), str(device))
        if key not in self._exchange:
            self._exchange[key] = make_exchange(P, device, self.group) if self.world > 1 else None
        return self._exchange[key]

    def forward(self, settings, means3D, shs, colors_precomp, opacities, scales, rotations, capacity=None):
        """Renders this rank's tiles (other tiles are zero).  Returns (color, normal, depth, opacity, state)."""
        from . import rasterizer as R
        H, W = int(settings.image_height), int(settings.image_width)
        mask = self.mask_for((H + 15) // 16, (W + 15) // 16, means3D.device)
        own = surfel_range(means3D.size(0), self.world, self.rank) if self.world > 1 else None
        color, normal, depth, opac, _active, _radii, st = R.forward_raw(
            settings, means3D, shs, colors_precomp, opacities, scales, rotations, mask, capacity=capacity, own_range=own)
        return color, normal, depth, opac, st

    def backward(self, st, means3D, shs, colors_precomp, scales, rotations, g_color, g_normal, g_depth, g_opac):
        """Reverse walk over this rank's tiles, the exchange of the screen-gradient rows, then the per-surfel backward
        on the owned range.  Returns (grads dict with full-size tensors whose rows outside [first, first+count) are
        zero, (first, count))."""
        from . import _lib, rasterizer as R
        lib = _lib.load()
        device = means3D.device
        P = means3D.size(0)
        Pp = padded_rows(P, self.world)
        f32 = dict(dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            stream = R._stream_ptr(device)
            ex = self.exchange_for(P, device)
            sg = torch.zeros((Pp, SCREEN_GRAD_STRIDE), **f32)
            gc, gn = R._f32c(g_color, device), R._f32c(g_normal, device)
            gd, go = R._f32c(g_depth, device), R._f32c(g_opac, device)
            _lib.check(lib.egs_backward_render(C.byref(st.frame), st.geom.data_ptr(), st.img.data_ptr(),
                                               st.bin.data_ptr(), st.cap, gc.data_ptr(), gn.data_ptr(), gd.data_ptr(),
                                               go.data_ptr(), sg.data_ptr(), _lib.EGS_BWD_GRADS_PREZEROED, stream),
                       "backward_render")
            first, count = surfel_range(P, self.world, self.rank)
            keep = sg
            if self.world > 1 and ex is not None:
                base = ex.exchange(st.geom, sg, stream)
            elif self.world > 1:
                keep = reduce_scatter_rows(sg, self.group)
                # the per-surfel kernel indexes the block by global surfel id: view the chunk at its global offset
                base = keep.data_ptr() - self.rank * (Pp // self.world) * SCREEN_GRAD_STRIDE * 4
            else:
                base = sg.data_ptr()
            use_sh = R._present(shs)
            M = st.frame.sh_coeffs
            out = {"means3D": torch.zeros((P, 3), **f32), "opacities": torch.zeros((P, 1), **f32),
                   "sh": torch.zeros((P, M, 3), **f32) if use_sh else None, "scales": torch.zeros((P, 3), **f32),
                   "rotations": torch.zeros((P, 4), **f32),
                   "colors_precomp": None if use_sh else torch.zeros((P, 3), **f32), "_keep": keep}
            means3D = R._f32c(means3D, device)
            shs_c = R._f32c(shs, device) if use_sh else None
            col_c = None if use_sh else R._f32c(colors_precomp, device)
            scales, rotations = R._f32c(scales, device), R._f32c(rotations, device)
            _lib.check(lib.egs_backward_surfels(C.byref(st.frame), first, count, means3D.data_ptr(), R._ptr(shs_c),
                                                R._ptr(col_c), scales.data_ptr(), rotations.data_ptr(),
                                                st.radii.data_ptr(), st.geom.data_ptr(), base,
                                                out["means3D"].data_ptr(), out["opacities"].data_ptr(),
                                                R._ptr(out["sh"]), out["scales"].data_ptr(),
                                                out["rotations"].data_ptr(), None, R._ptr(out["colors_precomp"]),
                                                None, stream), "backward_surfels")
            if self.world > 1 and ex is not None:
                ex.consumed()
        return out, (first, count)


# ------------------------------------------------------------------------------------------------- autograd level
class _ShardedRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, sh, colors_precomp, opacities, scales, rotations, raster_settings, sharder):
        color, normal, depth, opac, st = sharder.forward(raster_settings, means3D, sh, colors_precomp, opacities,
                                                         scales, rotations)
        ctx.sharder, ctx.state = sharder, st
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, sh)
        return color, normal, depth, opac

    @staticmethod
    def backward(ctx, g_color, g_normal, g_depth, g_opac):
        colors_precomp, means3D, scales, rotations, sh = ctx.saved_tensors
        g, _range = ctx.sharder.backward(ctx.state, means3D, sh, colors_precomp, scales, rotations, g_color, g_normal,
                                         g_depth, g_opac)
        return g["means3D"], g["sh"], g["colors_precomp"], g["opacities"], g["scales"], g["rotations"], None, None


class ShardedRasterizer(nn.Module):
    """`GaussianRasterizer` for one rank of a tile-sharded frame: same arguments as the reference's forward
    (diff_gaussian_rasterization/__init__.py:198-230) minus `tile_mask`, which the sharder owns.

    forward  -> (color, normal, depth, opacity): this rank's tiles, zeros elsewhere (`pixel_mask()` says which);
    backward -> the exchange runs inside; every rank receives the gradients of ITS surfel range
                (`owned_range(P)`), rows outside it are zero.  Summing a loss over `pixel_mask()` on every rank and
                all-reducing the scalar gives the single-GPU loss; concatenating the owned gradient rows gives the
                single-GPU gradients.
    """

    def __init__(self, raster_settings, sharder: Optional[ShardedSplat] = None):
        super().__init__()
        self.raster_settings = raster_settings
        self.sharder = sharder if sharder is not None else ShardedSplat()

    def pixel_mask(self) -> torch.Tensor:
        s = self.raster_settings
        H, W = int(s.image_height), int(s.image_width)
        m = self.sharder.mask_for((H + 15) // 16, (W + 15) // 16, s.viewmatrix.device)
        return m.bool().repeat_interleave(16, 0).repeat_interleave(16, 1)[:H, :W]

    def owned_range(self, P: int) -> Tuple[int, int]:
        return surfel_range(P, self.sharder.world, self.sharder.rank)

    def forward(self, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if scales is None or rotations is None:
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        return _ShardedRasterize.apply(means3D, empty if shs is None else shs,
                                       empty if colors_precomp is None else colors_precomp, opacities, scales,
                                       rotations, self.raster_settings, self.sharder)


# ------------------------------------------------------------------------------------------------- distributed mapper
class DistributedMapper:
    """`mapping.FusedMapper` over `world` GPUs: tile-sharded render + reverse walk, peer exchange of the touched
    screen-gradient rows, per-surfel backward + fused Adam on the owned surfel range, all-gather of the updated
    activated parameters.  Every rank holds a full `FrameBatchOptimizer` (parameters replicated; only
    the owned rows of the raw parameters and of the Adam state are ever updated locally, the activated ones are
    refreshed everywhere by the all-gather)."""

    def __init__(self, opt, width: int, height: int, capacity: int, sh_degree: int, group=None,
                 costs: Optional[torch.Tensor] = None, layout: str = "bands"):
        from . import _lib
        from .pipeline import SplatContext
        self.opt, self.group = opt, group
        self.world, self.rank = _world_rank(group)
        dev = opt.device
        P = opt.P
        self.first, self.count = surfel_range(P, self.world, self.rank)
        self.chunk = padded_rows(P, self.world) // self.world
        ty, tx = (height + 15) // 16, (width + 15) // 16
        self.tile_mask = tile_partition(ty, tx, self.world, self.rank, costs, layout).to(dev) if self.world > 1 else None
        self.exchange = make_exchange(P, dev, group) if self.world > 1 else None
        self.ctx = SplatContext(P, width, height, opt.M, capacity, device=dev, padded_rows=padded_rows(P, self.world),
                                own_range=(self.first, self.count) if self.world > 1 else None)
        f32 = dict(dtype=torch.float32, device=dev)
        self.terms = torch.zeros(_lib.EGM_TERMS, dtype=torch.float64, device=dev)
        self.g_color, self.g_normal = torch.empty((3, height, width), **f32), torch.empty((3, height, width), **f32)
        self.g_depth, self.g_opac = torch.empty((1, height, width), **f32), torch.zeros((1, height, width), **f32)
        # all-gather staging: one padded [world * chunk, 59 or so] buffer per activated tensor
        self._gather = {}

    def _all_gather_rows(self, name: str) -> None:
        o = self.opt
        all_gather_rows(getattr(o, name), self.first, self.count, self.chunk, self.world, self.group, self._gather,
                        full=o._full.get(name))

    def iterate(self, settings, frame_input, render_mask):
        from . import _lib, mapping as MP, rasterizer as R
        o, ctx = self.opt, self.ctx
        ctx.check_overflow()      # a previous frame of THIS rank's shard overflowed the binning capacity -> raise here
        ctx.set_camera(settings)
        with torch.no_grad():
            ctx.forward(o.xyz, o.shs, None, o.opacity, o.scales, o.rotations, self.tile_mask, watch_overflow=True)
            MP.loss_seed(ctx.color, ctx.depth, ctx.normal, frame_input["color_map"], frame_input.get("depth_map"),
                         frame_input.get("normal_map_c"), render_mask[0], render_mask[1], o.weights,
                         out=(self.terms, self.g_color, self.g_depth, self.g_normal), tile_mask=self.tile_mask)
            fused = o.can_fuse_sh
            if self.world == 1:
                ctx.backward_render(self.g_color, self.g_normal, self.g_depth, self.g_opac)
                if fused:
                    o.backward_surfels_and_step(ctx)
                else:
                    ctx.backward_surfels(o.xyz, o.shs, None, o.scales, o.rotations)
            else:
                ctx.backward_render(self.g_color, self.g_normal, self.g_depth, self.g_opac, prezeroed=self.exchange is not None)
                stream = R._stream_ptr(o.device)
                if self.exchange is not None:
                    base = self.exchange.exchange(ctx.geom, ctx.screen, stream)
                else:
                    mine = reduce_scatter_rows(ctx.screen, self.group)
                    base = mine.data_ptr() - self.rank * self.chunk * SCREEN_GRAD_STRIDE * 4
                if fused:
                    o.backward_surfels_and_step(ctx, self.first, self.count, screen_base=base)
                else:
                    ctx.backward_surfels(o.xyz, o.shs, None, o.scales, o.rotations, self.first, self.count,
                                         screen_base=base)
                if self.exchange is not None:
                    self.exchange.consumed()
            if not fused:
                o.step({"xyz": ctx.d_means, "shs": ctx.d_sh, "opacity": ctx.d_opac, "scales": ctx.d_scales,
                        "rotations": ctx.d_rots}, first=self.first, count=self.count)
            if self.world > 1:
                # the regulariser's norms and the image terms are sums over surfels / pixels: add the ranks' parts
                # (only the two slots this step wrote: the other norm slot already holds a global sum)
                slots = [(o.step_count + 1) & 1, 2]
                part = o.reg[slots]
                dist.all_reduce(part, group=self.group)
                o.reg[slots] = part
                part = self.terms[1:5].clone()
                dist.all_reduce(part, group=self.group)
                self.terms[1:5] = part
                for name in ("xyz", "shs", "opacity", "scales", "rotations"):
                    self._all_gather_rows(name)
            return o.loss_values(self.terms, frame_input.get("depth_map") is not None,
                                 frame_input.get("normal_map_c") is not None)

    def synchronize(self) -> None:
        """Waits for the iterations issued so far and raises if any of this rank's frames overflowed its capacity."""
        torch.cuda.synchronize(self.opt.device)
        self.ctx.check_overflow(block=True)
