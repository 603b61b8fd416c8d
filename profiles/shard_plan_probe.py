"""One-GPU emulation of one rank of a tile-sharded C3 frame: times the plan (projection + tile scan) and the render of
rank world//2 for world in {2,4,8}, both tile layouts, and reports how many surfels survive the candidate pass.
EGS_SHARD_TWO_PASS=0/1 forces the projection variant (read once per process).  Usage: python profiles/shard_plan_probe.py [C3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import parallel as par, pipeline as PL, rasterizer as R  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
dev = torch.device("cuda", 0)
scene, cams, grads, deg = bench.make_workload(name)
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
print("two_pass env:", os.environ.get("EGS_SHARD_TWO_PASS", "default"), "instances", I_full, flush=True)
WORLDS = [int(v) for v in os.environ.get("PROBE_WORLDS", "2,4,8").split(",")]
LAYOUTS = os.environ.get("PROBE_LAYOUTS", "rows,bands").split(",")
for world in WORLDS:
    for layout in LAYOUTS:
        rank = world // 2
        mask = par.tile_partition(ty, tx, world, rank, costs, layout).to(dev)
        first, count = par.surfel_range(P, world, rank)
        ctx = PL.SplatContext(P, W, H, 16, int(I_full / world * 1.6) + 4096, device=dev, own_range=(first, count))
        ctx.set_camera(s)
        times = {}
        ev = []

        def mark(k):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((k, e))

        for rep in range(25):
            ev.clear()
            mark("start")
            ctx.forward(params["xyz"], params["shs"], None, params["opacity"], params["scales"], params["rotations"],
                        mask, mark)
            torch.cuda.synchronize()
            if rep >= 5:
                for (k0, e0), (k1, e1) in zip(ev[:-1], ev[1:]):
                    times.setdefault(k1, []).append(e0.elapsed_time(e1))
        cnt = ctx.read_counters()
        cand = int(ctx.img[64:192].view(torch.int32).sum().cpu())
        kept = int((ctx.radii > 0).sum())
        print("world %d layout %-5s: plan %.4f ms render %.4f ms | instances %d, candidates %d (0 = single pass), radii>0 %d of %d"
              % (world, layout, np.mean(times["plan"]), np.mean(times["render"]), cnt[0], cand, kept, P), flush=True)
        del ctx
