"""Short driver for ncu: a few forward+backward steps of the UNMODIFIED reference rasterizer (oracle/_ref, compiled for
sm_100a by oracle/build_ref.sh) on the bench workload, through its own `_C` entry points.
Usage: python profiles/prof_reference.py [workload=C3] [steps=3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_loader  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ref = ref_loader.load()
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(name)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
p = {k: t(scene[k]) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
W, H = cams[0].width, cams[0].height
ty, tx = cams[0].tiles
mask = torch.ones((ty, tx), dtype=torch.int32, device=dev)
empty = torch.Tensor([])
bg = t(np.zeros(3, np.float32))
for i in range(steps):
    c, g = cams[i % len(cams)], grads[i % len(grads)]
    view, proj, campos = t(c.viewmatrix), t(c.projmatrix), t(c.campos)
    out = ref._C.rasterize_gaussians(bg, p["xyz"], empty, p["opacity"], p["scales"], p["rotations"], 1.0, empty, view, proj,
                                     mask, c.tanfovx, c.tanfovy, H, W, c.cx, c.cy, p["shs"], deg, campos, False, False)
    (I, tile_num, color, normal, depth, opac, active, radii, geomB, binB, imgB, tile_indices) = out
    ref._C.rasterize_gaussians_backward(tile_indices, tile_num, bg, p["xyz"], radii, empty, p["scales"], p["rotations"], 1.0,
                                        empty, view, proj, c.tanfovx, c.tanfovy, t(g["color"]), t(g["normal"]), t(g["depth"]),
                                        t(g["opacity"]), p["shs"], deg, campos, geomB, I, binB, imgB, False)
torch.cuda.synchronize()
print("reference: instances", int(I))
