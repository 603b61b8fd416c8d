"""Short driver for ncu: a few fused mapping iterations (eggfusion_b200.mapping.FusedMapper) of a synthetic workload.
Usage: python profiles/prof_mapping.py [workload=C3] [iters=3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import mapping as MP  # noqa: E402
from eggfusion_b200 import rasterizer as R  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(name)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
W, H = cams[0].width, cams[0].height
bg = t(np.zeros(3, np.float32))
settings = [E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, t(c.viewmatrix), t(c.projmatrix), deg,
                                            t(c.campos), False, False, c.cx, c.cy) for c in cams]
raw, frames = bench.mapping_inputs(scene, cams, dev)
opt = MP.FrameBatchOptimizer(raw, MP.LrParams(**bench.MAP_LR), MP.MappingWeights(**bench.MAP_WEIGHTS))
out = R.forward_raw(settings[0], opt.xyz, opt.shs, torch.Tensor([]), opt.opacity, opt.scales, opt.rotations, None)
I = out[6].num_rendered
del out
fm = MP.FusedMapper(opt, W, H, int(I * 1.1) + 4096, deg)
for i in range(iters):
    loss = fm.iterate(settings[i % len(settings)], *frames[i % len(frames)])
torch.cuda.synchronize()
print("instances", I, "loss", loss.cpu().numpy())
