"""Where the end-to-end time of the public API goes: forward+backward through GaussianRasterizer with given pixel
gradients / with the torch loss, with and without the per-step loss.item() read-back, and the CPU enqueue cost.
Usage: python profiles/e2e_probe.py [exact|auto]"""
import sys, time, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, eggfusion_b200 as E
from eggfusion_b200 import rasterizer as R
R.config.capacity = sys.argv[1] if len(sys.argv) > 1 else "exact"
print("capacity policy:", R.config.capacity)
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(sys.argv[2] if len(sys.argv) > 2 else "C3")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(scene[k]) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
c = cams[0]; W, H = c.width, c.height
bg = t(np.zeros(3, np.float32))
s = E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, t(c.viewmatrix), t(c.projmatrix), deg, t(c.campos), False, False, c.cx, c.cy)
tc = torch.rand(3, H, W, device=dev); td = torch.rand(1, H, W, device=dev) + 1
gc = torch.rand(3, H, W, device=dev); gn = torch.rand(3, H, W, device=dev); gd = torch.rand(1, H, W, device=dev)
def step(mode):
    color, normal, depth, opac, _a, _r = E.GaussianRasterizer(s)(means3D=leaf["xyz"], opacities=leaf["opacity"], shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"], tile_mask=None)
    if mode.startswith("loss"):
        loss = (color - tc).abs().mean() + (depth - td).abs().mean() + 0.1 * (1 - normal[2]).mean()
        loss.backward()
    else:
        torch.autograd.backward([color, normal, depth], [gc, gn, gd])
        loss = None
    for v in leaf.values(): v.grad = None
    if mode.endswith("item"):
        (loss if loss is not None else color[0, 0, 0]).item()
for mode in ["grad", "grad_item", "loss", "loss_item"]:
    for i in range(5): step(mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for i in range(50): step(mode)
    b.record(); torch.cuda.synchronize()
    print(mode, "gpu ms/step %.3f" % (a.elapsed_time(b) / 50), "wall %.3f" % ((time.perf_counter() - t0) * 1e3 / 50))
# CPU-side cost of one step without waiting for the GPU
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): step("grad")
t1 = time.perf_counter(); torch.cuda.synchronize()
print("cpu enqueue ms/step %.3f" % ((t1 - t0) * 1e3 / 20))
