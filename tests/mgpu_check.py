"""Multi-GPU correctness check (run under torchrun on N GPUs of one box, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py [workload] [iterations]

Three levels, each against a single-GPU run of the same inputs on rank 0:
  1. raw      -- parallel.ShardedSplat: every rank renders its tile shard, the screen-gradient rows are exchanged (peer
                 memory, or NCCL reduce-scatter with EGS_EXCHANGE=nccl), every rank runs the per-surfel backward on its
                 surfel range.  Images bit-identical, gradients <= 1e-5.
  2. autograd -- parallel.ShardedRasterizer + a torch loss over the rank's pixels + loss.backward(): the all-reduced loss
                 and the concatenated owned gradient rows equal the single-GPU GaussianRasterizer ones.
  3. mapper   -- parallel.DistributedMapper: `iterations` fused mapping iterations (render, loss, backward, Adam on the
                 owned range, all-gather of the parameters): parameters on EVERY rank equal the single-GPU FusedMapper's
                 <= 1e-6 (relative to the parameter scale).
Prints one line `MGPU_CHECK OK ...` per level or raises.  tests/test_mgpu_gpu.py launches it when >= 2 GPUs are visible."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import eggfusion_b200 as E  # noqa: E402
from eggfusion_b200 import mapping as MP, parallel as par, rasterizer as R, synthetic as syn  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def gather_rows(t, first, count, P, world, dev):
    chunk = par.padded_rows(P, world) // world
    v = t.reshape(P, -1)
    mine = torch.zeros((chunk, v.shape[1]), device=dev)
    mine[:count] = v[first:first + count]
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts)[:P]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
    mode = "peer" if par.make_exchange(16, dev) is not None else "nccl"

    # ---- 1. raw level, two steps (the second one exercises the alternate exchange block and the self-cleaning)
    # uniform costs -> contiguous row-major tile runs of equal length (the default "bands" layout, cut mid-row)
    uniform = torch.ones(((H + 15) // 16) * ((W + 15) // 16), dtype=torch.float64)
    sh = par.ShardedSplat(costs=uniform)
    for step in range(2):
        color, normal, depth, opacity, st = sh.forward(s, means, shs, empty, opac, scales, rots)
        grads, (first, count) = sh.backward(st, means, shs, empty, scales, rots, *gt)
    for img in (color, normal, depth, opacity):
        dist.all_reduce(img)
    full = {k: gather_rows(grads[k], first, count, P, world, dev) for k in ("means3D", "opacities", "sh", "scales", "rotations")}
    if rank == 0:
        c1, n1, d1, o1, _a, _r, st1 = R.forward_raw(s, means, shs, empty, opac, scales, rots, None)
        g1 = R.backward_raw(st1, means, shs, empty, scales, rots, *gt)
        errs = {"color": rel(color, c1), "normal": rel(normal, n1), "depth": rel(depth, d1), "opacity": rel(opacity, o1)}
        for k in full:
            errs["d_" + k] = rel(full[k], g1[k].reshape(P, -1))
        assert torch.equal(color, c1) and torch.equal(depth, d1), "sharded images must be bit-identical"
        assert max(errs.values()) <= 1e-5, errs
        print("MGPU_CHECK OK raw world=%d exchange=%s workload=%s instances=%d " % (world, mode, name, st1.num_rendered) +
              " ".join("%s=%.1e" % kv for kv in errs.items()), flush=True)
    dist.barrier()

    # ---- 2. autograd level
    leaf = {k: v.clone().requires_grad_(True) for k, v in (("xyz", means), ("opacity", opac), ("shs", shs),
                                                           ("scales", scales), ("rotations", rots))}
    tc = torch.rand(3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    rast = par.ShardedRasterizer(s, sh)
    px = rast.pixel_mask()
    color, normal, depth, opacity = rast(means3D=leaf["xyz"], opacities=leaf["opacity"], shs=leaf["shs"],
                                         scales=leaf["scales"], rotations=leaf["rotations"])
    n_px = float(H * W)
    loss = (((color - tc).abs().sum(0) + depth[0] + 0.1 * (1 - normal[2]) + opacity[0] ** 2) * px).sum() / n_px
    loss.backward()
    total = loss.detach().clone()
    dist.all_reduce(total)
    first, count = rast.owned_range(P)
    full = {k: gather_rows(leaf[k].grad, first, count, P, world, dev) for k in leaf}
    outside = sum(float(leaf[k].grad[:first].abs().sum() + leaf[k].grad[first + count:].abs().sum()) for k in leaf)
    assert outside == 0.0, "rows outside the owned range must be zero"
    if rank == 0:
        ref = {k: v.detach().clone().requires_grad_(True) for k, v in leaf.items()}
        c, n, d, o, _a, _r = E.GaussianRasterizer(s)(means3D=ref["xyz"], opacities=ref["opacity"], shs=ref["shs"],
                                                     scales=ref["scales"], rotations=ref["rotations"])
        l1 = ((c - tc).abs().sum(0) + d[0] + 0.1 * (1 - n[2]) + o[0] ** 2).sum() / n_px
        l1.backward()
        errs = {"loss": abs(float(total) - float(l1)) / abs(float(l1))}
        for k in full:
            errs["d_" + k] = rel(full[k], ref[k].grad.reshape(P, -1))
        assert max(errs.values()) <= 1e-5, errs
        print("MGPU_CHECK OK autograd world=%d exchange=%s " % (world, mode) + " ".join("%s=%.1e" % kv for kv in errs.items()),
              flush=True)
    dist.barrier()
    del leaf, full, color, normal, depth, opacity, grads
    torch.cuda.empty_cache()

    # ---- 3. distributed mapper
    scene, cams, _grads, deg_b = bench.make_workload(name)
    raw, frames = bench.mapping_inputs(scene, cams, dev)
    settings = [E.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, t(np.zeros(3, np.float32)), 1.0, t(c.viewmatrix),
                                                t(c.projmatrix), deg_b, t(c.campos), False, False, c.cx, c.cy) for c in cams]
    out = R.forward_raw(settings[0], t(scene["xyz"]), t(scene["shs"]), empty, t(scene["opacity"]), t(scene["scales"]),
                        t(scene["rotations"]), None)
    cap = int(out[6].num_rendered * 1.3) + 4096
    del out
    mopt = MP.FrameBatchOptimizer(raw, MP.LrParams(**bench.MAP_LR), MP.MappingWeights(**bench.MAP_WEIGHTS),
                                  padded_rows=par.padded_rows(P, world))
    dm = par.DistributedMapper(mopt, W, H, cap, deg_b, costs=uniform)
    losses = []
    for i in range(iters):
        losses.append(dm.iterate(settings[i % len(settings)], *frames[i % len(frames)]).clone())
    torch.cuda.synchronize()
    assert dm.ctx.read_counters()[2] == 0
    if rank == 0:
        ropt = MP.FrameBatchOptimizer(raw, MP.LrParams(**bench.MAP_LR), MP.MappingWeights(**bench.MAP_WEIGHTS))
        fm = MP.FusedMapper(ropt, W, H, cap, deg_b)
        ref_losses = [fm.iterate(settings[i % len(settings)], *frames[i % len(frames)]).clone() for i in range(iters)]
        want = {k: getattr(ropt, k).clone() for k in ("xyz", "shs", "opacity", "scales", "rotations")}
    else:
        want = {k: torch.empty_like(getattr(mopt, k)) for k in ("xyz", "shs", "opacity", "scales", "rotations")}
    errs = {}
    for k, v in want.items():          # every rank checks ITS replica against rank 0's single-GPU result
        dist.broadcast(v, 0)
        errs[k] = rel(getattr(mopt, k), v)
    worst = torch.tensor([max(errs.values())], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        lerr = max(abs(float(a[0]) - float(b[0])) / abs(float(b[0])) for a, b in zip(losses, ref_losses))
        moved = rel(want["xyz"], t(scene["xyz"]))
        assert float(worst) <= 1e-6, (errs, float(worst))
        assert lerr <= 1e-5, lerr
        assert moved > 0, "the optimiser did not move the parameters"
        print("MGPU_CHECK OK mapper world=%d exchange=%s iterations=%d worst_param_err(all ranks)=%.1e loss_err=%.1e "
              "last_loss=%.6f" % (world, mode, iters, float(worst), lerr, float(losses[-1][0])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
