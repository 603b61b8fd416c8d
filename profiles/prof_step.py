"""Short driver for ncu: a few forward+backward steps of a synthetic workload through the C-ABI
(eggfusion_b200.pipeline.SplatContext).  Usage: python profiles/prof_step.py [workload=C3] [steps=3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import rasterizer as R  # noqa: E402
from eggfusion_b200.pipeline import SplatContext  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(name)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(scene[k]) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
P, M = scene["xyz"].shape[0], scene["shs"].shape[1]
W, H = cams[0].width, cams[0].height
bg = t(np.zeros(3, np.float32))
settings = [E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, t(c.viewmatrix), t(c.projmatrix), deg,
                                            t(c.campos), False, False, c.cx, c.cy) for c in cams]
pix = [tuple(t(g[k]) for k in ("color", "normal", "depth", "opacity")) for g in grads]
out = R.forward_raw(settings[0], params["xyz"], params["shs"], torch.Tensor([]), params["opacity"], params["scales"],
                    params["rotations"], None)
I = out[6].num_rendered
del out
ctx = SplatContext(P, W, H, M, int(I * 1.1) + 4096, device=dev)
for i in range(steps):
    ctx.set_camera(settings[i % len(settings)])
    ctx.step(params, pix[i % len(pix)])
torch.cuda.synchronize()
print("instances", I, "counters", ctx.read_counters())
