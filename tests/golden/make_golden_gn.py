"""Golden vectors for the fused dense-tracker Gauss-Newton step (include/eggtrack.h egt_gn_*),
produced by the REFERENCE'S OWN python code (/root/reference/src/core/optimizer.py: projective_transform,
icp_optimization, rgb_optimization, update_transform) imported unmodified and run with torch on the CPU in the build
container (only `easydict`, absent here and unused by these functions, is stubbed).

    python tests/golden/make_golden_gn.py        # needs /root/reference; writes tests/golden/gn_*.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")

CASES = {
    # name: (W, H, seed, translation scale, rotation scale (rad))
    "gn_96x72": (96, 72, 21, 0.004, 0.004),
    "gn_61x45_far": (61, 45, 22, 0.03, 0.03),     # large motion: many pixels leave the image / fail the thresholds
}
ANGLE_THRES, DIST_THRES = 20.0, 0.1               # configs/replica/base.yaml:31-32


def _rot(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def case_inputs(name):
    """One pyramid level of a model (prev) and a frame (curr) + a pose, float32 numpy (PyraImageCUDA layouts)."""
    W, H, seed, ts, rs = CASES[name]
    r = np.random.default_rng(seed)
    fx = fy = 0.9 * W
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")

    def maps(phase):
        z = 2.0 + 0.3 * np.sin(xx / W * 5 + phase) * np.cos(yy / H * 4) + 0.002 * r.standard_normal((H, W))
        v = np.stack([(xx - cx) / fx * z, (yy - cy) / fy * z, z], -1)
        dx = np.gradient(v, axis=1)
        dy = np.gradient(v, axis=0)
        nrm = np.cross(dy, dx)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        I = 0.5 + 0.4 * np.sin(xx * 0.9 + phase) * np.cos(yy * 0.7) + 0.05 * r.standard_normal((H, W))
        gx_, gy_ = np.gradient(I, axis=1) * 8, np.gradient(I, axis=0) * 8
        grad = np.stack([gx_, gy_, np.sqrt(gx_ ** 2 + gy_ ** 2 + 1e-6)], -1)
        mask = r.uniform(0, 1, (H, W, 1)) < 0.93
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        return {"vertex": f(v), "normal": f(nrm), "intensity": f(I[..., None]), "grad": f(grad), "mask": mask,
                "disp": f(1.0 / (z[..., None] + 1e-6))}
    model, frame = maps(0.0), maps(0.02)
    frame["normal"][5, 7] = np.nan                 # invalid normals must be masked out, not propagated
    frame["vertex"][9, 11] = np.nan
    T = np.eye(4)
    T[:3, :3] = _rot(rs * r.standard_normal(3))
    T[:3, 3] = ts * r.standard_normal(3)
    dx = np.concatenate([0.01 * r.standard_normal(3), 0.02 * r.standard_normal(3)]).astype(np.float32)
    return model, frame, np.float32([fx, fy, cx, cy]), T.astype(np.float32), dx


def run_reference(name):
    import torch
    sys.path.insert(0, REF)
    stub = types.ModuleType("easydict")
    stub.EasyDict = dict
    sys.modules.setdefault("easydict", stub)
    import src.core.optimizer as O

    model, frame, intr, T, dx = case_inputs(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    pyr = lambda m: types.SimpleNamespace(
        vertex_pyramid=[t(m["vertex"])], normal_pyramid=[t(m["normal"])], mask_pyramid=[t(m["mask"])],
        intensity_pyramid=[t(m["intensity"])], grad_pyramid=[t(m["grad"])], disp_pyramid=[t(m["disp"])],
        intrinsic_pyramid=[t(intr)])
    pm, pf = pyr(model), pyr(frame)
    Tt = t(T)
    coords, Jc = O.projective_transform(Tt, pm.disp_pyramid[0], pm.intrinsic_pyramid[0])
    A_icp, b_icp, n_icp = O.icp_optimization(pm, pf, 0, Tt, coords, ANGLE_THRES, DIST_THRES)
    A_rgb, b_rgb, n_rgb = O.rgb_optimization(pm, pf, 0, coords, Jc)
    T2 = O.update_transform(Tt.clone(), t(dx))
    T3 = O.update_transform(Tt.clone(), t(dx * 1e-5))     # small-angle branch of so3_to_SO3
    return {"coords": coords.numpy(), "A_icp": A_icp.numpy(), "b_icp": b_icp.numpy().reshape(-1), "n_icp": np.int64(n_icp),
            "A_rgb": A_rgb.numpy(), "b_rgb": b_rgb.numpy().reshape(-1), "n_rgb": np.int64(n_rgb),
            "T_updated": T2.numpy(), "T_updated_small": T3.numpy()}


if __name__ == "__main__":
    for name in CASES:
        res = run_reference(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, "icp", int(res["n_icp"]), "rgb", int(res["n_rgb"]), "of", res["coords"].shape[0] * res["coords"].shape[1])
