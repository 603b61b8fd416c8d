"""Forward with and without backward state (EGS_FWD_NO_SAVE) at C3: time per forward and peak memory.  Usage: python profiles/fwd_only_time.py"""
import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, eggfusion_b200 as E
from eggfusion_b200 import rasterizer as R
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload("C3")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
p = {k: t(scene[k]) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
c = cams[0]
s = E.GaussianRasterizationSettings(c.height, c.width, c.tanfovx, c.tanfovy, t(np.zeros(3, np.float32)), 1.0, t(c.viewmatrix), t(c.projmatrix), deg, t(c.campos), False, False, c.cx, c.cy)
R.config.capacity = "auto"
empty = torch.Tensor([])
for save in (True, False):
    f = lambda: R.forward_raw(s, p["xyz"], p["shs"], empty, p["opacity"], p["scales"], p["rotations"], None, save=save)
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): f()
    b.record(); torch.cuda.synchronize()
    print("save" if save else "forward-only", "ms/forward %.3f" % (a.elapsed_time(b) / 50), "peak MB %.0f" % (torch.cuda.max_memory_allocated() / 1e6))
    torch.cuda.reset_peak_memory_stats()
