"""Golden vectors for the mapping-iteration glue (include/eggmap.h), produced by the REFERENCE'S OWN
python code running on the CPU in the build container:

  * `Mapper.compute_loss`, `Mapper.total_params` and `check_nan` are cut out of /root/reference/src/core/mapper.py by
    AST (the module itself cannot be imported: easydict, cv2-less deps and the CUDA extensions are absent) and
    executed unmodified;
  * `GaussianSurfels` (activations, get_normal, parametrize) is imported from the reference with a stub for the absent
    `plyfile`; its tensors are placed on the CPU and `torch.zeros(..., device="cuda")` inside build_rotation is
    redirected to the CPU;
  * the optimiser is `torch.optim.Adam(surfels.parametrize(lr), lr=0.0)` exactly as mapper.py:338.

The rasterizer is replaced by a linear surrogate  sum(activated_param * G)  with fixed random G: autograd then delivers
to the activations exactly the gradients `egs_backward_surfels` would (= G), which is all the glue sees of the render.

    python tests/golden/make_golden_mapping.py        # needs /root/reference; writes tests/golden/mapping_*.npz
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")

CASES = {
    # name: (P, H, W, sh_coeffs, iterations, weights (color, depth, normal, reg, reg_n), seed)
    "mapping_deg3": (257, 40, 56, 16, 3, (1.0, 1.0, 1.0, 10.0, 1.0), 11),
    "mapping_deg0_noreg": (130, 24, 32, 1, 2, (1.0, 0.5, 2.0, 0.0, 1.0), 12),
}
LR = dict(position_lr=1e-5 * 100, feature_lr=1e-3, opacity_lr=1e-5 * 100, scaling_lr=5e-4, rotation_lr=1e-4 * 10)


def case_inputs(name):
    """Seeded inputs (numpy, float32).  Shared by the golden generator and the tests."""
    P, H, W, M, K, weights, seed = CASES[name]
    r = np.random.default_rng(seed)
    f = lambda *s: r.standard_normal(s).astype(np.float32)
    raw = {
        "xyz": f(P, 3),
        "features_dc": 0.5 * f(P, 1, 3),
        "features_rest": 0.1 * f(P, M - 1, 3),
        "scaling": np.concatenate([np.log(r.uniform(0.01, 0.05, (P, 2))).astype(np.float32),
                                   np.full((P, 1), -1.0e10, np.float32)], axis=1),
        "rotation": (f(P, 4) * r.uniform(0.2, 3.0, (P, 1)).astype(np.float32)),
        "opacity": f(P, 1) * 2,
    }
    raw["scaling"][5, 2] = np.log(0.2)          # a surfel whose thinnest axis is not z ...
    raw["scaling"][5, 0] = np.log(0.001)        # ... (argmin = 0)
    raw["scaling"][6, 1] = np.log(0.0005)
    raw["scaling"][6, 2] = np.log(0.3)          # argmin = 1
    its = []
    for _ in range(K):
        est_n = f(3, H, W)
        est_n[:, :3, :] = 0.0                   # uncovered pixels: zero normal -> eps-clamped norm in cosine_similarity
        ref_n = f(H, W, 3)
        ref_n /= np.linalg.norm(ref_n, axis=-1, keepdims=True)
        est_c = r.uniform(0, 1, (3, H, W)).astype(np.float32)
        ref_c = r.uniform(0, 1, (H, W, 3)).astype(np.float32)
        ref_c[4, :8, :] = est_c[:, 4, :8].T     # exact ties: sign(0) = 0
        est_n[:, 6, :8] = ref_n[6, :8, :].T * 2.0   # parallel: cos = 1 -> clamped, zero gradient
        sparse = (r.uniform(0, 1, (P, 1)) < 0.6).astype(np.float32)
        its.append({
            "est_color": est_c, "est_depth": r.uniform(0.5, 3, (1, H, W)).astype(np.float32), "est_normal": est_n,
            "ref_color": ref_c, "ref_depth": r.uniform(0.5, 3, (H, W, 1)).astype(np.float32), "ref_normal": ref_n,
            "rgb_mask": r.uniform(0, 1, (H, W)) < 0.8, "geo_mask": r.uniform(0, 1, (H, W)) < 0.7,
            "G_xyz": 1e-3 * f(P, 3) * sparse, "G_shs": 1e-3 * f(P, M, 3) * sparse[:, :, None],
            "G_opacity": 1e-3 * f(P, 1) * sparse, "G_scales": 1e-2 * f(P, 3) * sparse, "G_rot": 1e-3 * f(P, 4) * sparse,
        })
    return raw, its, weights, LR


def _extract(path, names):
    """Source of the named top-level functions / class methods of a python file, unmodified."""
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in out:
            node.decorator_list = []
            out[node.name] = ast.get_source_segment(src, node)
    return out


def run_reference(name):
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, REF)
    stub = types.ModuleType("plyfile")
    stub.PlyData = stub.PlyElement = object
    sys.modules.setdefault("plyfile", stub)
    from src.core.gaussian_surfels import GaussianSurfels
    import src.core.utils as ref_utils

    class _TorchCPU:   # torch.zeros((..), device="cuda") inside build_rotation -> CPU (no GPU in the build container)
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def zeros(*a, **kw):
            kw.pop("device", None)
            return torch.zeros(*a, **kw)
    ref_utils.torch = _TorchCPU()

    fns = _extract(os.path.join(REF, "src/core/mapper.py"), {"compute_loss", "total_params", "check_nan"})
    ns = {"torch": torch, "F": F}
    for k in ("check_nan", "compute_loss", "total_params"):
        import textwrap
        exec(textwrap.dedent(fns[k]), ns)

    raw, its, (cw, dw, nw, rw, rwn), lr = case_inputs(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    surf = object.__new__(GaussianSurfels)
    surf.setup_functions()
    surf._xyz, surf._features_dc, surf._features_rest = t(raw["xyz"]), t(raw["features_dc"]), t(raw["features_rest"])
    surf._scaling, surf._rotation, surf._opacity = t(raw["scaling"]), t(raw["rotation"]), t(raw["opacity"])
    mapper = types.SimpleNamespace(surfels0=surf, color_weight=cw, depth_weight=dw, normal_weight=nw, reg_weight=rw,
                                   reg_weight_n=rwn)
    cfg = types.SimpleNamespace(**lr)
    optimizer = torch.optim.Adam(surf.parametrize(cfg), lr=0.0)                      # mapper.py:338
    geo = {"position": surf.get_xyz.detach().clone(), "normal": surf.get_normal.detach().clone()}   # mapper.py:342-345
    out = {"normal0": geo["normal"].numpy().copy()}
    for k, it in enumerate(its):
        tp = ns["total_params"](mapper)
        est = {n: t(it["est_" + n]).requires_grad_(True) for n in ("color", "depth", "normal")}
        frame_input = {"color_map": t(it["ref_color"]), "depth_map": t(it["ref_depth"]), "normal_map_c": t(it["ref_normal"])}
        loss = ns["compute_loss"](mapper, est, frame_input, (t(it["rgb_mask"]), t(it["geo_mask"])), geo)
        surrogate = ((tp["xyz"] * t(it["G_xyz"])).sum() + (tp["shs"] * t(it["G_shs"])).sum()
                     + (tp["opacity"] * t(it["G_opacity"])).sum() + (tp["scales"] * t(it["G_scales"])).sum()
                     + (tp["rotations"] * t(it["G_rot"])).sum())
        (loss + surrogate).backward()
        out[f"loss_{k}"] = np.float32(loss.item())
        for n in est:
            out[f"seed_{n}_{k}"] = est[n].grad.numpy().copy()
        for n in ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity"):
            out[f"grad_{n}_{k}"] = getattr(surf, "_" + n).grad.numpy().copy()
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
        for n in ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity"):
            prm = getattr(surf, "_" + n)
            out[f"param_{n}_{k}"] = prm.detach().numpy().copy()
            out[f"m_{n}_{k}"] = optimizer.state[prm]["exp_avg"].numpy().copy()
            out[f"v_{n}_{k}"] = optimizer.state[prm]["exp_avg_sq"].numpy().copy()
        tp = ns["total_params"](mapper)
        for n in ("opacity", "scales", "rotations", "normal"):
            out[f"act_{n}_{k}"] = tp[n].detach().numpy().copy()
    return out


if __name__ == "__main__":
    for name in CASES:
        res = run_reference(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, {k: float(v) for k, v in res.items() if k.startswith("loss_")},
              "%.0f kB" % (os.path.getsize(os.path.join(HERE, name + ".npz")) / 1e3))
