"""How selective is the forward's per-(instance, 8x4 block) hit test?  Runs one forward of a workload, exports
records / point list / tile ranges and recomputes with torch (analysis only, not a product path):
  box    -- pairs the extent-box test of k_render_forward passes
  exact  -- pairs where the alpha >= 1/255 ellipse really reaches a pixel of the block (ignoring transmittance stop)
Usage: python profiles/hit_stats.py [workload=C3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import rasterizer as R  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(name)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(scene[k]) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
P = scene["xyz"].shape[0]
c = cams[0]
W, H = c.width, c.height
bg = t(np.zeros(3, np.float32))
s = E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, t(c.viewmatrix), t(c.projmatrix), deg,
                                    t(c.campos), False, False, c.cx, c.cy)
out = R.forward_raw(s, params["xyz"], params["shs"], torch.Tensor([]), params["opacity"], params["scales"],
                    params["rotations"], None)
st = out[6]
d = R.debug_export(st, P, W, H)
gx = (W + 15) // 16
ranges = d["ranges"].long()
cnt = ranges[:, 1] - ranges[:, 0]
tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0], device=dev), cnt)
# tile offsets are an exclusive scan in tile-id order, so list position -> tile is a repeat_interleave
pl = d["point_list"].long()
I = pl.numel()
tile_pos = tile_of
assert tile_pos.numel() == I
rec = d["records"][pl]                     # [I,16]
x, y = rec[:, 0], rec[:, 1]
ext = rec[:, 2].view(torch.int32)
hx = (ext & 0xffff).float() * 0.125
hy = ((ext >> 16) & 0xffff).float() * 0.125
op = rec[:, 3]
cxx, cxy, cyy = rec[:, 4], rec[:, 5], rec[:, 6]
tx0 = (tile_pos % gx).float() * 16
ty0 = (tile_pos // gx).float() * 16
box = exact = 0
n_px_box = 0
for b in range(8):
    bx0 = tx0 + (b & 1) * 8
    by0 = ty0 + (b >> 1) * 4
    hit = ((x - (bx0 + 3.5)).abs() <= hx + 3.5) & ((y - (by0 + 1.5)).abs() <= hy + 1.5)
    any_px = torch.zeros_like(hit)
    for py in range(4):
        for px in range(8):
            dx = x - (bx0 + px)
            dy = y - (by0 + py)
            power = -0.5 * (cxx * dx * dx + cyy * dy * dy) - cxy * dx * dy
            alpha = torch.clamp(op * torch.exp(power), max=0.99)
            any_px |= (power <= 0) & (alpha >= 1.0 / 255.0)
    box += int(hit.sum())
    exact += int((any_px & hit).sum())
    missed = int((any_px & ~hit).sum())
    assert missed == 0, missed
print(f"{name}: instances {I}  box pairs {box} ({box / I:.2f}/instance)  exact pairs {exact} ({exact / I:.2f}/instance)"
      f"  false-hit share {(box - exact) / box:.3f}")
