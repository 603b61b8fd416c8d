"""Multi-GPU sharding of one render/optimise step (one process per GPU, torch.distributed for the plumbing).

The reference is single-GPU; its `tile_mask` argument (forward.cu:292-300, rasterizer_impl.cu:103-111) is the seam:
a surfel only emits instances for tiles whose mask is non-zero.  Scheme:

  * every rank holds the full surfel parameter set and runs the cheap per-surfel projection for all of it;
  * rank r bins / sorts / composites only its tile set (interleaved tile rows, or cost-balanced from the previous
    frame's per-tile list lengths) and runs the reverse walk over the same tiles -> a partial screen-space
    gradient block G_r[P][16];
  * ONE collective: reduce-scatter(sum) of G over ranks, so rank r receives the summed rows of its surfel range;
  * rank r runs the per-surfel backward for its range only -> gradient shards [first, first+count).

The exchange moves 64 B per surfel instead of the 236 B of parameter gradients at SH degree 3.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import SCREEN_GRAD_STRIDE


def tile_partition(tiles_y: int, tiles_x: int, world: int, rank: int,
                   costs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32 [tiles_y, tiles_x] mask of the tiles rank `rank` renders.  Masks of all ranks are disjoint and cover
    the grid.  Without `costs`: tile rows are dealt round-robin.  With `costs` (per-tile list lengths, any
    device): rows are assigned greedily (longest-processing-time first) to balance the summed cost."""
    mask = torch.zeros((tiles_y, tiles_x), dtype=torch.int32)
    if costs is None:
        mask[rank::world, :] = 1
        return mask
    row_cost = costs.detach().to("cpu", torch.float64).reshape(tiles_y, tiles_x).sum(dim=1)
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


def padded_rows(P: int, world: int) -> int:
    """Rows of the exchanged block: P rounded up so every rank owns the same number of rows."""
    return (P + world - 1) // world * world


def surfel_range(P: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of the surfel rows rank `rank` owns after the reduce-scatter."""
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


class ShardedSplat:
    """One rank's share of a tile-sharded forward + backward.  CUDA only."""

    def __init__(self, group=None, costs: Optional[torch.Tensor] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.costs = costs
        self._mask_cache = {}

    def mask_for(self, tiles_y: int, tiles_x: int, device) -> torch.Tensor:
        key = (tiles_y, tiles_x, str(device))
        if key not in self._mask_cache:
            self._mask_cache[key] = tile_partition(tiles_y, tiles_x, self.world, self.rank, self.costs).to(device)
        return self._mask_cache[key]

    def forward(self, settings, means3D, shs, colors_precomp, opacities, scales, rotations, capacity=None):
        """Renders this rank's tiles (other tiles are zero).  Returns (color, normal, depth, opacity, state)."""
        from . import rasterizer as R
        H, W = int(settings.image_height), int(settings.image_width)
        mask = self.mask_for((H + 15) // 16, (W + 15) // 16, means3D.device)
        color, normal, depth, opac, _active, _radii, st = R.forward_raw(
            settings, means3D, shs, colors_precomp, opacities, scales, rotations, mask, capacity=capacity)
        return color, normal, depth, opac, st

    def backward(self, st, means3D, shs, colors_precomp, scales, rotations, g_color, g_normal, g_depth, g_opac):
        """Reverse walk over this rank's tiles, ONE reduce-scatter of the screen-gradient block, then the
        per-surfel backward on the owned range.  Returns (grads dict with full-size tensors whose rows outside
        [first, first+count) are unspecified, (first, count))."""
        import ctypes as C
        from . import _lib, rasterizer as R
        lib = _lib.load()
        device = means3D.device
        P = means3D.size(0)
        Pp = padded_rows(P, self.world)
        f32 = dict(dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            stream = R._stream_ptr(device)
            sg = torch.empty((Pp, SCREEN_GRAD_STRIDE), **f32)
            if Pp > P:
                sg[P:].zero_()
            gc, gn = R._f32c(g_color, device), R._f32c(g_normal, device)
            gd, go = R._f32c(g_depth, device), R._f32c(g_opac, device)
            _lib.check(lib.egs_backward_render(C.byref(st.frame), st.geom.data_ptr(), st.img.data_ptr(),
                                               st.bin.data_ptr(), st.cap, gc.data_ptr(), gn.data_ptr(), gd.data_ptr(),
                                               go.data_ptr(), sg.data_ptr(), 0, stream), "backward_render")
            first, count = surfel_range(P, self.world, self.rank)
            if self.world > 1:
                mine = reduce_scatter_rows(sg, self.group)
                # the per-surfel kernel indexes the block by global surfel id: view the chunk at its global offset
                chunk = Pp // self.world
                base = mine.data_ptr() - self.rank * chunk * SCREEN_GRAD_STRIDE * 4
            else:
                mine, base = sg, sg.data_ptr()
            use_sh = R._present(shs)
            M = st.frame.sh_coeffs
            out = {"means3D": torch.empty((P, 3), **f32), "opacities": torch.empty((P, 1), **f32),
                   "sh": torch.empty((P, M, 3), **f32) if use_sh else None, "scales": torch.empty((P, 3), **f32),
                   "rotations": torch.empty((P, 4), **f32),
                   "colors_precomp": None if use_sh else torch.empty((P, 3), **f32), "_keep": mine}
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
        return out, (first, count)
