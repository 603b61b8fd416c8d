"""Runs the REFERENCE's own, unmodified Python loop (oracle/_ref/egg: a byte copy of /root/reference/src + configs made by
oracle/build_ref.sh) on a synthetic RGB-D sequence, on top of one of two native back ends:

    --arm ours        sys.path gets eggfusion_b200/dropin: `diff_gaussian_rasterization` and `cuda_tracking_ext` resolve to
                      this repository's drop-ins (libeggsplat.so)
    --arm reference   `diff_gaussian_rasterization` = oracle/_ref (the reference's own CUDA build), `cuda_tracking_ext` =
                      the reference's own build (tests/shims_ref: its CPU Eigen solve replaced by torch.linalg.lstsq)

What runs is main.py:39-66's loop: EGGFusion(cfg); per frame Frame.init_from_dataset -> EGGFusion.reconstruct =
Tracker.tracking (tracker.py:124-252) -> preprocess -> Mapping.mapping (mapper.py:180-378: surfel fusion, sampling,
Renderer.render, frame_batch_optimization with compute_loss + torch Adam) -> postprocess (system.py:44-125).
Configuration: the reference's configs/replica/office0.yaml or configs/tum/fr1_desk.yaml through its own 3-level merge,
with `Tracking.use_sparse: False` (ORB-SLAM2 is un-buildable offline) and `Dataset.preload: False`; the datasets are
absent offline, so frames come from eggfusion_b200.synthetic.make_rgbd_frame at the dataset's calibration.
Missing pure-python dependencies (omegaconf, easydict, plyfile, open3d, matplotlib, ...) are tests/shims stand-ins.

Writes an .npz: estimated and ground-truth poses, the model map rendered at the last frame, surfel count, per-frame
wall times.  Usage: python tests/ref_loop.py --arm ours --config replica --frames 12 --out /tmp/ours.npz
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EGG = os.path.join(ROOT, "oracle", "_ref", "egg")

CONFIGS = {"replica": "configs/replica/office0.yaml", "tum": "configs/tum/fr1_desk.yaml"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["ours", "reference"], required=True)
    ap.add_argument("--config", choices=sorted(CONFIGS), default="replica")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    args.out = os.path.abspath(args.out)
    if not os.path.isdir(os.path.join(EGG, "src")):
        raise SystemExit("oracle/_ref/egg is missing: run oracle/build_ref.sh where /root/reference exists")

    # import resolution: stand-ins for the absent pure-python packages, then the native back end of the arm
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
    if args.arm == "ours":
        sys.path.insert(0, os.path.join(ROOT, "eggfusion_b200", "dropin"))
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        sys.path.insert(0, os.path.join(ROOT, "tests", "shims_ref"))     # in front of the raw .so in oracle/_ref
    sys.path.insert(0, ROOT)
    sys.path.insert(0, EGG)                       # `src.*`, exactly as main.py's working directory provides it
    os.chdir(EGG)                                 # the configs name each other by relative path

    import torch
    from easydict import EasyDict as edict
    from omegaconf import OmegaConf

    scene_config = OmegaConf.load(CONFIGS[args.config])                  # main.py:15-19
    cfg = OmegaConf.merge(OmegaConf.load(scene_config.base_config), OmegaConf.load(scene_config.data_config), scene_config)
    cfg.Tracking.use_sparse = False
    cfg.Dataset.preload = False
    cfg.System.save_dir = "/tmp/egg_loop_%s_%s" % (args.arm, args.config)
    os.makedirs(cfg.System.save_dir, exist_ok=True)

    import diff_gaussian_rasterization as dgr
    import cuda_tracking_ext as cte
    from src.system import EGGFusion
    from src.utils.frame import Frame
    from src.utils.camera_utils import focal2fov, getProjectionMatrix_v2
    from eggfusion_b200 import synthetic as syn

    backend = {"rasterizer": os.path.relpath(dgr.__file__, ROOT), "tracking": os.path.relpath(cte.__file__, ROOT)}
    expect = "eggfusion_b200/dropin" if args.arm == "ours" else "oracle/_ref"
    assert backend["rasterizer"].startswith(expect), backend
    assert backend["tracking"].startswith("eggfusion_b200/dropin" if args.arm == "ours" else "tests/shims_ref"), backend
    print("arm", args.arm, "back ends:", backend, flush=True)

    calib = cfg.Dataset.Calibration

    class SyntheticRGBD:
        """Duck-types what Frame.init_from_dataset needs of RGBDDataset (dataset.py:29-115): `params` and __getitem__."""

        def __init__(self):
            fovx, fovy = focal2fov(calib.fx, calib.width), focal2fov(calib.fy, calib.height)
            self.params = edict({"fx": calib.fx, "fy": calib.fy, "cx": calib.cx, "cy": calib.cy, "width": calib.width,
                                 "height": calib.height, "fovx": fovx, "fovy": fovy,
                                 "projection_matrix": getProjectionMatrix_v2(znear=0.01, zfar=100.0, fovX=fovx,
                                                                             fovY=fovy).transpose(0, 1),
                                 "depth_scale": calib.depth_scale})
            self.frames = [syn.make_rgbd_frame(i, calib.width, calib.height, calib.fx, calib.fy, calib.cx, calib.cy,
                                               calib.depth_scale) for i in range(args.frames)]

        def __len__(self):
            return len(self.frames)

        def __getitem__(self, i):
            return self.frames[i]

    dataset = SyntheticRGBD()
    torch.manual_seed(20251201)       # the reference samples new surfels with torch.rand (mapper.py:446-492): same draws in both arms
    np.random.seed(20251201)
    ef = EGGFusion(cfg)
    est, ref, wall = [], [], []
    for fid in range(len(dataset)):                                    # main.py:52-63
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        curr_frame = Frame.init_from_dataset(dataset, fid, cfg.Dataset.preload)
        ef.reconstruct(curr_frame)
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
        est.append(curr_frame.c2w_matrix().cpu().numpy())
        ref.append(curr_frame.c2w_matrix(gt=True).cpu().numpy())
        torch.cuda.empty_cache()
    with torch.no_grad():
        rendered = ef.mapper.get_render_output(curr_frame)
    est, ref = np.stack(est), np.stack(ref)
    ate = float(np.sqrt(np.mean(np.sum((est[:, :3, 3] - ref[:, :3, 3]) ** 2, axis=1))))
    n_surfels = int(ef.mapper.surfels0._xyz.shape[0])
    # steady-state frame rate: frame 0 runs the 20-iteration initialisation
    steady = wall[1:] if len(wall) > 1 else wall
    print("arm %s config %s: %d frames, %d surfels, ATE rmse %.5f m, %.1f ms/frame steady (%.2f frames/s), first frame %.1f ms"
          % (args.arm, args.config, len(wall), n_surfels, ate, 1e3 * float(np.mean(steady)), 1.0 / float(np.mean(steady)),
             1e3 * wall[0]), flush=True)
    np.savez_compressed(args.out, est=est, ref=ref, wall=np.asarray(wall), ate=ate, n_surfels=n_surfels,
                        color=rendered["render_color"].detach().cpu().numpy(),
                        depth=rendered["render_depth"].detach().cpu().numpy(),
                        opacity=rendered["render_opacity"].detach().cpu().numpy())


if __name__ == "__main__":
    main()
