"""How often are consecutive entries of a block's hit list disjoint in their pixel masks?  (Would two hits fit in one
phase-1 pass of the backward?)  Analysis only.  Usage: python profiles/hit_pair_stats.py [workload=C3]"""
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
torch.cuda.synchronize()
cap = st.cap
al = lambda v: (v + 255) // 256 * 256
off_hits = al(al(8 * cap) + 4 * cap)
hits = st.bin[off_hits: off_hits + 64 * cap].view(torch.int32).view(-1, 2).cpu().numpy()
tiles = ((W + 15) // 16) * ((H + 15) // 16)
# img workspace: counters(256) | tile_count | tile_offset | tile_cursor | tile_list | hit_count
o = 256 + al(4 * tiles) + al(4 * (tiles + 1)) + al(4 * tiles) + al(4 * tiles)
hit_count = st.img[o: o + 32 * tiles].view(torch.int32).cpu().numpy()
ranges = d["ranges"].cpu().numpy().astype(np.int64)
tot = pairs = disjoint = merged = 0
pop = []
rng = np.random.default_rng(0)
for tile in rng.choice(tiles, size=600, replace=False):
    a, b = ranges[tile]
    n = b - a
    for blk in range(8):
        cnt = hit_count[8 * tile + blk]
        if cnt == 0:
            continue
        m = hits[8 * a + blk * n: 8 * a + blk * n + cnt, 1].astype(np.uint32)[::-1]   # the backward's order
        tot += cnt
        pop.append(np.mean([bin(int(x)).count("1") for x in m]))
        pairs += cnt - 1
        disjoint += int(np.sum((m[:-1] & m[1:]) == 0))
        i = 0
        while i + 1 < cnt:          # greedy merging of neighbours
            if (m[i] & m[i + 1]) == 0:
                merged += 1
                i += 2
            else:
                i += 1
print("%s: hits %d  mean active pixels %.1f  consecutive pairs disjoint %.3f  passes saved by greedy merging %.3f"
      % (name, tot, float(np.mean(pop)), disjoint / max(pairs, 1), merged / max(tot, 1)))
