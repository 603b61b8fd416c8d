"""One-GPU emulation of the sender side of the screen-gradient exchange (egs_push_rows) and the owner's fold: rank
world//2 of `world` ranks, contiguous tile runs, every "peer" inbox mapped to a local buffer (so NVLink is out of the
picture: what remains is the kernel's own structure).  Usage: python profiles/push_probe.py [world]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import _lib, parallel as par, pipeline as PL, rasterizer as R  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload("C3")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
P = scene["xyz"].shape[0]
W, H = cams[0].width, cams[0].height
params = {k: t(scene[k]) for k in ("xyz", "opacity", "shs", "scales", "rotations")}
c = cams[0]
s = E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, t(np.zeros(3, np.float32)), 1.0, t(c.viewmatrix),
                                    t(c.projmatrix), deg, t(c.campos), False, False, c.cx, c.cy)
empty = torch.Tensor([])
out = R.forward_raw(s, params["xyz"], params["shs"], empty, params["opacity"], params["scales"], params["rotations"], None)
rg = R.debug_export(out[6], P, W, H)["ranges"].double()
costs = rg[:, 1] - rg[:, 0]
I_full = out[6].num_rendered
del out
ty, tx = c.tiles
rank = world // 2
mask = par.tile_partition(ty, tx, world, rank, costs, "bands").to(dev)
rows = par.padded_rows(P, world)
chunk = rows // world
first, count = par.surfel_range(P, world, rank)
ctx = PL.SplatContext(P, W, H, 16, int(I_full / world * 1.6) + 4096, device=dev, padded_rows=rows, own_range=(first, count))
ctx.set_camera(s)
pix = tuple(t(grads[0][k]) for k in ("color", "normal", "depth", "opacity"))
lib = _lib.load()
inbox = torch.zeros((world * chunk * 16,), dtype=torch.float32, device=dev)
header = torch.zeros((world,), dtype=torch.int32, device=dev)
sent = torch.zeros((world,), dtype=torch.int32, device=dev)
inboxes = torch.tensor([inbox.data_ptr()] * world, dtype=torch.int64, device=dev)
headers = torch.tensor([header.data_ptr()] * world, dtype=torch.int64, device=dev)
block = torch.zeros((chunk, 16), dtype=torch.float32, device=dev)
stream = R._stream_ptr(dev)
tp, tf = [], []
for rep in range(12):
    ctx.forward(params["xyz"], params["shs"], None, params["opacity"], params["scales"], params["rotations"], mask)
    ctx.backward_render(*pix, prezeroed=rep > 0)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    _lib.check(lib.egs_push_rows(P, chunk, world, rank, ctx.geom.data_ptr(), ctx.screen.data_ptr(), sent.data_ptr(),
                                 inboxes.data_ptr(), headers.data_ptr(), stream), "push")
    e[1].record()
    _lib.check(lib.egs_fold_inbox(chunk, world, first, inbox.data_ptr(), header.data_ptr(), block.data_ptr(), stream), "fold")
    e[2].record()
    torch.cuda.synchronize()
    if rep >= 2:
        tp.append(e[0].elapsed_time(e[1]))
        tf.append(e[1].elapsed_time(e[2]))
touched = int((ctx.geom.view(torch.uint8)[:0].numel() == 0))
print("world %d: push %.4f ms fold %.4f ms | rows sent per owner (all land in one local header): %s, screen block zero after push: %s"
      % (world, np.mean(tp), np.mean(tf), header.tolist(), bool((ctx.screen == 0).all())), flush=True)
