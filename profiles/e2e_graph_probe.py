"""Can the public API (GaussianRasterizer + a torch loss + loss.backward()) be recorded into a CUDA graph and replayed?
Checks replayed losses / gradients against eager ones and times both.  Usage: python profiles/e2e_graph_probe.py [C3]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, eggfusion_b200 as E
from eggfusion_b200 import rasterizer as R

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(name)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
leaf = {k: t(scene[k]).requires_grad_(True) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
c = cams[0]; W, H = c.width, c.height
bg = t(np.zeros(3, np.float32))
view, proj, campos = t(c.viewmatrix), t(c.projmatrix), t(c.campos)       # static tensors: refreshed by copy_ before a replay
tc = torch.rand(3, H, W, device=dev); td = torch.rand(1, H, W, device=dev) + 1
s = E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, view, proj, deg, campos, False, False, c.cx, c.cy)

def step():
    color, normal, depth, opac, _a, _r = E.GaussianRasterizer(s)(means3D=leaf["xyz"], opacities=leaf["opacity"], shs=leaf["shs"],
                                                                 scales=leaf["scales"], rotations=leaf["rotations"], tile_mask=None)
    loss = (color - tc).abs().mean() + (depth - td).abs().mean() + 0.1 * (1 - normal[2]).mean()
    loss.backward()
    return loss

R.config.capacity = "auto"
for i in range(3):
    for v in leaf.values(): v.grad = None
    l_eager = float(step())       # keep no reference to the eager autograd graph (its AccumulateGrad nodes live on the default stream)
torch.cuda.synchronize()
g_eager = {k: v.grad.clone() for k, v in leaf.items()}
for v in leaf.values(): v.grad = None
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(2):
        for v in leaf.values(): v.grad = None
        step()
torch.cuda.current_stream().wait_stream(side)
for v in leaf.values(): v.grad = None
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    l_static = step()
g.replay(); torch.cuda.synchronize()
print("counters of captured forwards:", R.check_captured())
print("loss eager %.8f graph %.8f" % (l_eager, float(l_static)))
for k in leaf:
    d = float((leaf[k].grad - g_eager[k]).abs().max() / (g_eager[k].abs().max() + 1e-30))
    print("  grad", k, "rel diff %.2e" % d)
# second camera through the same graph
c2 = cams[1]
view.copy_(t(c2.viewmatrix)); proj.copy_(t(c2.projmatrix)); campos.copy_(t(c2.campos))
g.replay(); torch.cuda.synchronize(); l2g = float(l_static); g2 = leaf["xyz"].grad.clone()
for v in leaf.values(): v.grad = None
l2e = float(step()); torch.cuda.synchronize()
print("camera 2: loss eager %.8f graph %.8f, grad xyz rel diff %.2e" % (l2e, l2g, float((leaf["xyz"].grad - g2).abs().max() / g2.abs().max())))
def timeit(fn, n=100):
    for i in range(5): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
def eager_item():
    for v in leaf.values(): v.grad = None
    return float(step())
def graph_item():
    g.replay(); return l_static.item()
print("eager + item: %.3f ms/step; graph replay + item: %.3f ms/step" % (timeit(eager_item), timeit(graph_item)))
