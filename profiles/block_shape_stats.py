"""Which warp-block shape should the compositing kernels use?  For every (instance, block) pair of one C3 forward:
how many pairs pass the extent-box test (= pairs a warp visits) for blocks of 8x4 (1 px/lane), 8x8 and 16x4 (2 px/lane)
and 16x8 (4 px/lane), and how many (pixel, splat) pairs pass the alpha test (transmittance stop ignored).
Analysis only (torch), not a product path.   Usage: python profiles/block_shape_stats.py [workload=C3]"""
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
s = E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, t(np.zeros(3, np.float32)), 1.0, t(c.viewmatrix),
                                    t(c.projmatrix), deg, t(c.campos), False, False, c.cx, c.cy)
out = R.forward_raw(s, params["xyz"], params["shs"], torch.Tensor([]), params["opacity"], params["scales"],
                    params["rotations"], None)
st = out[6]
d = R.debug_export(st, P, W, H)
gx = (W + 15) // 16
ranges = d["ranges"].long()
cnt = ranges[:, 1] - ranges[:, 0]
tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0], device=dev), cnt)
pl = d["point_list"].long()
I = pl.numel()
rec = d["records"][pl]
x, y = rec[:, 0], rec[:, 1]
ext = rec[:, 2].view(torch.int32)
hx = (ext & 0xffff).float() * 0.125
hy = ((ext >> 16) & 0xffff).float() * 0.125
op = rec[:, 3]
cxx, cxy, cyy = rec[:, 4], rec[:, 5], rec[:, 6]
tx0 = (tile_of % gx).float() * 16
ty0 = (tile_of // gx).float() * 16
# per-pixel alpha pass over the tile: [I] x 256 in chunks
alpha_px = torch.zeros((I, 16, 16), dtype=torch.bool, device=dev)
for py in range(16):
    for px in range(16):
        dx = x - (tx0 + px)
        dy = y - (ty0 + py)
        power = -0.5 * (cxx * dx * dx + cyy * dy * dy) - cxy * dx * dy
        alpha = torch.clamp(op * torch.exp(power), max=0.99)
        alpha_px[:, py, px] = (power <= 0) & (alpha >= 1.0 / 255.0)
pairs = int(alpha_px.sum())
print(f"{name}: instances {I}; (pixel, splat) pairs passing the alpha test {pairs} ({pairs / I:.1f}/instance)")
for bw, bh in ((8, 4), (8, 8), (16, 4), (16, 8), (4, 4), (16, 16)):
    box = exact = 0
    for by in range(0, 16, bh):
        for bx in range(0, 16, bw):
            hit = ((x - (tx0 + bx + (bw - 1) / 2)).abs() <= hx + (bw - 1) / 2) & \
                  ((y - (ty0 + by + (bh - 1) / 2)).abs() <= hy + (bh - 1) / 2)
            anyp = alpha_px[:, by:by + bh, bx:bx + bw].reshape(I, -1).any(dim=1)
            box += int(hit.sum())
            exact += int((anyp & hit).sum())
    print(f"  block {bw:2d}x{bh:2d}: box pairs {box} ({box / I:.2f}/instance), exact {exact} ({exact / I:.2f}/instance), "
          f"lane utilisation {pairs / (box * bw * bh):.3f}, px-slots {box * bw * bh / 1e6:.0f} M")
