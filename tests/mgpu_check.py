"""Multi-GPU correctness check (run under torchrun on N GPUs of one box, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py [workload]

Every rank renders its tile shard, the screen-gradient block is reduce-scattered, every rank runs the per-surfel
backward on its own surfel range; rank 0 gathers the shards and compares images and gradients with a single-GPU
run of the same frame.  Prints one line `MGPU_CHECK OK ...` or raises."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import parallel as par, rasterizer as R, synthetic as syn  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P, W, H, L, deg = syn.CONFIGS[name]
    cam = syn.default_camera(W, H, syn.look_from((0.03, -0.02, 0.02), 0.02, -0.01))
    sc = syn.make_scene(P, syn.default_camera(W, H), layers=L, sh_degree=deg)
    g = syn.make_pixel_grads(cam, with_opacity=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    s = E.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, t(np.zeros(3, np.float32)), 1.0,
                                        t(cam.viewmatrix), t(cam.projmatrix), deg, t(cam.campos), False, False,
                                        cam.cx, cam.cy)
    empty = torch.Tensor([])
    means, shs, opac = t(sc["xyz"]), t(sc["shs"]), t(sc["opacity"])
    scales, rots = t(sc["scales"]), t(sc["rotations"])
    gt = [t(g[k]) for k in ("color", "normal", "depth", "opacity")]

    sh = par.ShardedSplat()
    color, normal, depth, opacity, st = sh.forward(s, means, shs, empty, opac, scales, rots)
    grads, (first, count) = sh.backward(st, means, shs, empty, scales, rots, *gt)
    # assemble: images are disjoint per rank -> sum; gradient shards -> all_gather of padded chunks
    for img in (color, normal, depth, opacity):
        dist.all_reduce(img)
    chunk = par.padded_rows(P, world) // world
    full = {}
    for k in ("means3D", "opacities", "sh", "scales", "rotations"):
        v = grads[k].reshape(P, -1)
        mine = torch.zeros((chunk, v.shape[1]), device=dev)
        mine[:count] = v[first:first + count]
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        full[k] = torch.cat(parts)[:P]
    if rank == 0:
        c1, n1, d1, o1, _a, _r, st1 = R.forward_raw(s, means, shs, empty, opac, scales, rots, None)
        g1 = R.backward_raw(st1, means, shs, empty, scales, rots, *gt)
        rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
        errs = {"color": rel(color, c1), "normal": rel(normal, n1), "depth": rel(depth, d1), "opacity": rel(opacity, o1)}
        for k in full:
            errs["d_" + k] = rel(full[k], g1[k].reshape(P, -1))
        assert torch.equal(color, c1) and torch.equal(depth, d1), "sharded images must be bit-identical"
        assert max(errs.values()) <= 1e-5, errs
        print("MGPU_CHECK OK world=%d workload=%s instances=%d " % (world, name, st1.num_rendered) +
              " ".join("%s=%.1e" % kv for kv in errs.items()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
