"""Generates tests/golden/*.npz with the UNMODIFIED reference rasterizer (oracle/_ref, built by
oracle/build_ref.sh from /root/reference) on a CUDA GPU.  Run on the B200 box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy the .npz files into tests/golden/ and commit them.  Each file holds the seeded inputs' identity
(config name, seed, camera pose id), the reference's images, gradients and every index artefact decoded from
its geometry / binning / image workspaces.  The oracle and the CUDA path are both tested against these files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from eggfusion_b200 import synthetic as syn  # noqa: E402
from oracle import ref_loader  # noqa: E402

# name -> (P, W, H, layers, sh_degree, pose, with_opacity_grad, bg, tile_mask_kind)
CASES = {
    "c1_identity": (10_000, 256, 256, 2, 3, None, False, (0, 0, 0), "ones"),
    "c1_posed_bg": (10_000, 256, 256, 2, 3, ((0.15, -0.1, 0.05), 0.12, -0.07), True, (0.2, 0.5, 0.7), "ones"),
    "small_deg0_ragged": (3_000, 200, 136, 3, 0, ((-0.05, 0.08, -0.1), -0.1, 0.05), True, (0, 0, 0), "checker"),
    "small_deg1": (2_000, 160, 96, 2, 1, None, False, (0, 0, 0), "ones"),
    "small_deg2": (2_000, 160, 96, 2, 2, None, False, (0, 0, 0), "ones"),
}


def case_inputs(name):
    P, W, H, L, deg, pose, with_op, bg, mask_kind = CASES[name]
    base = syn.default_camera(W, H)
    sc = syn.make_scene(P, base, layers=L, sh_degree=deg)
    cam = base if pose is None else syn.default_camera(W, H, syn.look_from(*pose))
    g = syn.make_pixel_grads(cam, with_opacity=with_op)
    ty, tx = cam.tiles
    mask = np.ones((ty, tx), np.int32)
    if mask_kind == "checker":
        mask[::2, 1::2] = 0
        mask[1::2, ::2] = 0
    return cam, sc, g, np.asarray(bg, np.float32), mask, deg


def run_reference(name, device="cuda"):
    ref = ref_loader.load()
    cam, sc, g, bg, mask, deg = case_inputs(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    P = sc["xyz"].shape[0]
    M = sc["shs"].shape[1]
    settings = ref.GaussianRasterizationSettings(
        image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t(bg),
        scale_modifier=1.0, viewmatrix=t(cam.viewmatrix), projmatrix=t(cam.projmatrix), sh_degree=deg,
        campos=t(cam.campos), prefiltered=False, debug=False, cx=cam.cx, cy=cam.cy)
    means, shs, opac, scales, rots = t(sc["xyz"]), t(sc["shs"]), t(sc["opacity"]), t(sc["scales"]), t(sc["rotations"])
    empty = torch.Tensor([])
    args = (settings.bg, means, empty, opac, scales, rots, settings.scale_modifier, empty, settings.viewmatrix,
            settings.projmatrix, t(mask), settings.tanfovx, settings.tanfovy, settings.image_height,
            settings.image_width, settings.cx, settings.cy, shs, settings.sh_degree, settings.campos, False, False)
    (num_rendered, num_tile, color, normal, depth, opac_img, active, radii, geomB, binB, imgB,
     tile_indices) = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    bargs = (tile_indices, num_tile, settings.bg, means, radii, empty, scales, rots, settings.scale_modifier, empty,
             settings.viewmatrix, settings.projmatrix, settings.tanfovx, settings.tanfovy, t(g["color"]),
             t(g["normal"]), t(g["depth"]), t(g["opacity"]), shs, settings.sh_degree, settings.campos, geomB,
             num_rendered, binB, imgB, False)
    (d_means2D, d_colors, d_opacity, d_means3D, d_cov3D, d_sh, d_scales, d_rots) = \
        ref._C.rasterize_gaussians_backward(*bargs)
    torch.cuda.synchronize()
    geom = ref_loader.decode_geom(geomB, P)
    img = ref_loader.decode_img(imgB, cam.width * cam.height)
    binn = ref_loader.decode_binning(binB, num_rendered)
    tiles = cam.tiles[0] * cam.tiles[1]
    vis = radii.cpu().numpy() > 0
    out = {
        "num_rendered": np.int64(num_rendered), "tile_num": np.int64(num_tile),
        "color": color.cpu().numpy(), "normal_img": normal.cpu().numpy(), "depth": depth.cpu().numpy(),
        "opacity": opac_img.cpu().numpy(), "active_mask": active.cpu().numpy(), "radii": radii.cpu().numpy(),
        "tile_indices": tile_indices.cpu().numpy()[:tiles].copy(),
        "ranges": img["ranges"][:tiles].copy(), "n_contrib": img["n_contrib"].copy(),
        "final_T": img["accum_alpha"].copy(), "final_D": img["accum_depth"].copy(),
        "point_list": binn["point_list"].copy(), "point_list_keys": binn["point_list_keys"].copy(),
        "tiles_touched": geom["tiles_touched"].copy(),
        # per-surfel state is only defined for visible surfels (the reference leaves the rest uninitialised)
        "vis_index": np.nonzero(vis)[0].astype(np.int32),
        "means2D": geom["means2D"][vis], "depths": geom["depths"][vis], "cov3D": geom["cov3D"][vis],
        "conic_opacity": geom["conic_opacity"][vis], "rgb": geom["rgb"][vis], "normal": geom["normal"][vis],
        "Jinv": geom["Jinv"][vis], "clamped": geom["clamped"][vis],
        "dL_dmeans2D": d_means2D.cpu().numpy(), "dL_dcolors": d_colors.cpu().numpy(),
        "dL_dopacity": d_opacity.cpu().numpy(), "dL_dmeans3D": d_means3D.cpu().numpy(),
        "dL_dcov3D": d_cov3D.cpu().numpy(), "dL_dsh": d_sh.cpu().numpy(), "dL_dscales": d_scales.cpu().numpy(),
        "dL_drotations": d_rots.cpu().numpy(),
    }
    return out


FUSION_CASES = {
    # name: (P, W, H, pose)
    "fusion_320x240": (20_000, 320, 240, ((0.05, -0.02, 0.03), 0.04, -0.03)),
}


def fusion_inputs(name):
    P, W, H, pose = FUSION_CASES[name]
    cam = syn.default_camera(W, H, syn.look_from(*pose))
    return cam, syn.make_fusion_case(P, cam)


def run_reference_fusion(name, device="cuda"):
    """_C.project_surfels_to_frame + _C.preprocess_surfels of the unmodified reference."""
    ref = ref_loader.load()
    cam, fc = fusion_inputs(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    P = fc["points"].shape[0]
    pts, rot, s2 = t(fc["points"]), t(fc["rotations"]), t(fc["sigma2"])
    stable = t(fc["stable_mask"])
    intr, view, proj = t(fc["intrinsic"]), t(cam.viewmatrix), t(cam.projmatrix)
    imap, dbuf = ref.project_surfels_to_frame(pts, rot, stable, intr, view, proj, 1.0, cam.height, cam.width)
    torch.cuda.synchronize()
    z = lambda *shape, dt=torch.float32: torch.zeros(*shape, dtype=dt, device=device)
    inview, surface = z(P, dt=torch.bool), z(P, dt=torch.bool)
    ref.preprocess_surfels(pts, rot, z(P, 3), z(P, 3), z(P), z(P, dt=torch.int32), z(P, 6), s2, z(P, dt=torch.int32),
                           z(P, dt=torch.int32), stable, intr, view, proj, t(fc["frame_vmap"]), t(fc["frame_nmap"]),
                           z(cam.height, cam.width, 3), t(fc["frame_dmap"]), t(fc["frame_mask"]), imap, dbuf,
                           z(cam.height, cam.width, 3), z(cam.height, cam.width, 3),
                           z(cam.height, cam.width, dt=torch.bool), inview, surface, fc["fusion_dist_thres"],
                           fc["alpha_p"], fc["alpha_n"])
    torch.cuda.synchronize()
    return {"index_map": imap.cpu().numpy(), "depth_buffer": dbuf.cpu().numpy(), "points": pts.cpu().numpy(),
            "rotations": rot.cpu().numpy(), "sigma2": s2.cpu().numpy(), "inview_mask": inview.cpu().numpy(),
            "surface_mask": surface.cpu().numpy()}


def tracking_inputs(seed=20251209, W=161, H=119):
    """Deterministic frame for the dense-tracking utilities (odd sizes exercise the borders)."""
    rng = np.random.default_rng(seed)
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    depth = 1.5 + 0.4 * np.sin(u / 17.0) * np.cos(v / 11.0) + 0.002 * rng.normal(size=(H, W))
    depth[(u > 100) & (v < 30)] += 0.8                      # a depth discontinuity
    depth[rng.uniform(size=(H, W)) < 0.02] = 0.0            # sensor holes
    gray = 0.5 + 0.3 * np.sin(u / 5.0 + v / 9.0) + 0.05 * rng.normal(size=(H, W))
    rgb = np.stack([gray, 0.8 * gray + 0.1, 1.0 - gray], axis=-1)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return {"depth": f32(depth), "gray": f32(gray), "rgb": f32(rgb), "intr": (140.0, 138.0, 80.3, 59.1)}


def load_ref_tracking():
    """The reference's cuda_tracking_ext built by oracle/build_ref.sh (Eigen stubbed: kernels only)."""
    import glob
    import importlib.util as iu
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cuda_tracking_ext*.so"))
    if not so:
        return None
    spec = iu.spec_from_file_location("cuda_tracking_ext", so[0])
    mod = iu.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference_tracking(device="cuda"):
    ext = load_ref_tracking()
    ti = tracking_inputs()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    depth, gray, rgb = t(ti["depth"]), t(ti["gray"]), t(ti["rgb"])
    H, W = depth.shape
    fx, fy, cx, cy = ti["intr"]
    out = {}
    o = torch.zeros_like(depth)
    ext.bilateral_filter_cuda(depth, o, W, H, 13, 0.03, 4.5)
    out["bilateral"] = o.cpu().numpy()
    o = torch.zeros_like(rgb)
    ext.gaussian_filter_cuda(rgb, o, W, H, 3, 5, 1.5)
    out["gaussian"] = o.cpu().numpy()
    for name, img in (("down1", gray[..., None].contiguous()), ("down3", rgb)):
        o = torch.zeros(H // 2, W // 2, img.shape[2], device=device)
        ext.gaussian_downsample_cuda(img, o, W, H, img.shape[2])
        out[name] = o.cpu().numpy()
    gx, gy = torch.zeros_like(gray), torch.zeros_like(gray)
    ext.compute_gradients_cuda(gray, gx, gy, W, H)
    out["grad_x"], out["grad_y"] = gx.cpu().numpy(), gy.cpu().numpy()
    vm, nm = torch.zeros(H, W, 3, device=device), torch.zeros(H, W, 3, device=device)
    ext.compute_vertex_and_normal_cuda(depth, fx, fy, cx, cy, vm, nm)
    out["vertex"], out["normal"] = vm.cpu().numpy(), nm.cpu().numpy()
    return out


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name in CASES:
        out = run_reference(name)
        # second run: the float atomics of the reference backward are order-nondeterministic; record the spread
        out2 = run_reference(name)
        for k in ("dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations", "dL_dopacity"):
            out["rerun_absdiff_" + k] = np.float64(np.abs(out[k] - out2[k]).max())
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print(name, "I=%d tiles=%d vis=%d" % (out["num_rendered"], out["tile_num"], len(out["vis_index"])),
              "size=%.2f MB" % (os.path.getsize(os.path.join(outdir, name + ".npz")) / 1e6))
    if load_ref_tracking() is not None:
        np.savez_compressed(os.path.join(outdir, "tracking_161x119.npz"), **run_reference_tracking())
        print("tracking_161x119 written")
    for name in FUSION_CASES:
        out = run_reference_fusion(name)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print(name, "index hits %.3f inview %.3f surface %.3f" % ((out["index_map"] >= 0).mean(),
                                                                  out["inview_mask"].mean(), out["surface_mask"].mean()))


if __name__ == "__main__":
    main()
