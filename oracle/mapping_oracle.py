"""CPU oracle (numpy, float32) for the mapping-iteration glue of include/eggmap.h -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product
(eggfusion_b200/) never does.

Restates, operation by operation, what PyTorch computes for the reference's python code:
  loss_seed   Mapper.compute_loss image terms (/root/reference/src/core/mapper.py:381-426,437-438) and the gradients
              autograd returns for est_color / est_depth / est_normal
  activate    GaussianSurfels.get_opacity / get_scaling / get_rotation / get_normal
              (/root/reference/src/core/gaussian_surfels.py:345-393, core/utils.py:69-92) + nan_to_num of
              Mapper.total_params (mapper.py:565-585)
  adam_step   regulariser of compute_loss (mapper.py:427-435) + backward of the activations + torch.optim.Adam.step()
              for the parameter groups of GaussianSurfels.parametrize (gaussian_surfels.py:134-150)
PINNED: tests/golden/mapping_*.npz were produced by the reference's own functions (cut out of mapper.py by AST,
GaussianSurfels imported) with torch autograd + torch.optim.Adam on the CPU (tests/golden/make_golden_mapping.py);
tests/test_mapping_cpu.py checks this file against them.
"""
import numpy as np

F32 = np.float32
EPS_COS = F32(1e-8)
LO, HI = F32(-1 + 1e-6), F32(1 - 1e-6)


def _sign(x):
    return np.sign(x).astype(F32)


def cosdist(x1, x2, up):
    """|1 - clamp(F.cosine_similarity(x1, x2, dim=-1))| and up * d/dx2 of it (ATen cosine_similarity: norms clamped to eps
    outside autograd, linalg_vector_norm backward masked at 0).  x1, x2: [..., 3]."""
    n1t = np.sqrt((x1 * x1).sum(-1, keepdims=True, dtype=F32))
    n2t = np.sqrt((x2 * x2).sum(-1, keepdims=True, dtype=F32))
    n1, n2 = np.maximum(n1t, EPS_COS), np.maximum(n2t, EPS_COS)
    a, b = x1 / n1, x2 / n2
    c = (a * b).sum(-1, dtype=F32)
    cd = F32(1) - np.clip(c, LO, HI)
    g = (-up * _sign(cd) * ((c >= LO) & (c <= HI)))[..., None].astype(F32)
    y = g * a
    e = y / n2
    dn2 = -(y * (b / n2)).sum(-1, keepdims=True, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        e = e + np.where(n2t > 0, x2 * (dn2 / n2t), F32(0))
    return np.abs(cd), e.astype(F32)


def loss_seed(est_color, est_depth, est_normal, ref_color, ref_depth, ref_normal, rgb_mask, geo_mask, cw, dw, nw):
    """-> dict(count, color_loss, depth_loss, normal_loss, image_loss, seeds dL_dcolor [3,H,W], dL_ddepth [1,H,W],
    dL_dnormal [3,H,W])."""
    m = rgb_mask & geo_mask if geo_mask is not None else rgb_mask
    n = int(m.sum())
    ec, ed, en = est_color.transpose(1, 2, 0), est_depth.transpose(1, 2, 0), est_normal.transpose(1, 2, 0)
    out = {"count": n}
    gc, gd, gn = np.zeros_like(ec), np.zeros_like(ed), np.zeros_like(en)
    x = ref_color - ec
    out["color_loss"] = F32(np.abs(x[m]).mean(dtype=np.float64)) if n else F32(np.nan)
    if n:
        gc[m] = (-_sign(x) * F32(cw / (3.0 * n)))[m]
    out["depth_loss"] = F32(0)
    if ref_depth is not None and dw > 0 and n:
        x = ref_depth - ed
        out["depth_loss"] = F32(np.abs(x[m]).mean(dtype=np.float64))
        gd[m] = (-_sign(x) * F32(dw / n))[m]
    out["normal_loss"] = F32(0)
    if ref_normal is not None and nw > 0 and n:
        cd, e = cosdist(ref_normal, en, F32(nw / n))
        out["normal_loss"] = F32(cd[m].mean(dtype=np.float64))
        gn[m] = e[m]
    out["image_loss"] = F32(cw * out["color_loss"] + dw * out["depth_loss"] + nw * out["normal_loss"])
    out["dL_dcolor"] = np.ascontiguousarray(gc.transpose(2, 0, 1))
    out["dL_ddepth"] = np.ascontiguousarray(gd.transpose(2, 0, 1))
    out["dL_dnormal"] = np.ascontiguousarray(gn.transpose(2, 0, 1))
    return out


def _normalize_quat(raw):
    nq = np.sqrt((raw * raw).sum(-1, keepdims=True, dtype=F32))
    den = np.maximum(nq, F32(1e-12))
    return raw / den, nq, den


def _column(q, k):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    cols = [
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)], -1),
        np.stack([2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x)], -1),
        np.stack([2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)], -1),
    ]
    return np.choose(k[:, None], cols).astype(F32)


def get_normal(qh, scales):
    """GaussianSurfels.get_normal: column argmin(scales) of build_rotation(qh) / (norm + 1e-8)."""
    k = np.argmin(scales, axis=1)
    nb = np.sqrt(qh[:, 0] * qh[:, 0] + qh[:, 1] * qh[:, 1] + qh[:, 2] * qh[:, 2] + qh[:, 3] * qh[:, 3])[:, None]
    q = qh / nb
    v = _column(q, k)
    mag = np.sqrt((v * v).sum(-1, keepdims=True, dtype=F32))
    return (v / (mag + F32(1e-8))).astype(F32), (k, nb, q, v, mag)


def _get_normal_bwd(qh, saved, dn):
    k, nb, q, v, mag = saved
    d = mag + F32(1e-8)
    dv = dn / d
    dmag = -(dn * ((v / d) / d)).sum(-1, keepdims=True, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        dv = dv + np.where(mag > 0, v * (dmag / mag), F32(0))
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    d0, d1, d2 = dv[:, 0], dv[:, 1], dv[:, 2]
    dq_k = [
        np.stack([2 * z * d1 - 2 * y * d2, 2 * y * d1 + 2 * z * d2, -4 * y * d0 + 2 * x * d1 - 2 * r * d2,
                  -4 * z * d0 + 2 * r * d1 + 2 * x * d2], -1),
        np.stack([-2 * z * d0 + 2 * x * d2, 2 * y * d0 - 4 * x * d1 + 2 * r * d2, 2 * x * d0 + 2 * z * d2,
                  -2 * r * d0 - 4 * z * d1 + 2 * y * d2], -1),
        np.stack([2 * y * d0 - 2 * x * d1, 2 * z * d0 - 2 * r * d1 - 4 * x * d2, 2 * r * d0 + 2 * z * d1 - 4 * y * d2,
                  2 * x * d0 + 2 * y * d1], -1),
    ]
    dq = np.choose(k[:, None], dq_k).astype(F32)
    dnb = -(dq * (q / nb)).sum(-1, keepdims=True, dtype=F32)
    return (dq / nb + dnb * (qh / nb)).astype(F32)


def activate(opacity_raw, scaling_raw, rotation_raw):
    """-> opacity, scales, rotations (nan_to_num'd), normal (GaussianSurfels.get_normal)."""
    with np.errstate(over="ignore"):
        opacity = (F32(1) / (F32(1) + np.exp(-opacity_raw))).astype(F32)
        scales = np.exp(scaling_raw).astype(F32)
    qh, _, _ = _normalize_quat(rotation_raw)
    normal, _ = get_normal(qh, scales)
    return opacity, scales, np.nan_to_num(qh, nan=1.0).astype(F32), normal


def adam_update(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (defaults) for one tensor; returns p, m, v."""
    m = (m + F32(1 - beta1) * (g - m)).astype(F32)
    v = (v * F32(beta2) + F32(1 - beta2) * g * g).astype(F32)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / F32(bc2 ** 0.5) + F32(eps)
    return (p + F32(-(lr / bc1)) * (m / denom)).astype(F32), m, v


def adam_step(raw, grads, state, lr, step, reg_weight=0.0, reg_weight_n=0.0, pos0=None, normal0=None):
    """One optimiser step.  raw: dict xyz, features_dc, features_rest, scaling, rotation, opacity (updated copies are
    returned); grads: dict G_xyz, G_shs, G_opacity, G_scales, G_rot (w.r.t. activated parameters); state: dict
    name -> (m, v).  Returns (new_raw, new_state, raw_grads, reg_loss)."""
    P = raw["xyz"].shape[0]
    opacity, scales, _, _ = activate(raw["opacity"], raw["scaling"], raw["rotation"])
    qh, nq, den = _normalize_quat(raw["rotation"])
    g = {
        "xyz": grads["G_xyz"].copy(),
        "features_dc": grads["G_shs"][:, :1].copy(),
        "features_rest": grads["G_shs"][:, 1:].copy(),
        "opacity": (grads["G_opacity"] * (F32(1) - opacity) * opacity).astype(F32),
        "scaling": (grads["G_scales"] * scales).astype(F32),
    }
    dqh = np.where(np.isfinite(qh), grads["G_rot"], F32(0)).astype(F32)
    reg_loss = F32(0)
    if reg_weight > 0:
        diff = pos0 - raw["xyz"]
        nrm = np.sqrt((diff.astype(np.float64) ** 2).sum())
        if nrm > 0:
            g["xyz"] = (g["xyz"] + (raw["xyz"] - pos0) * F32(reg_weight / nrm)).astype(F32)
        n, saved = get_normal(qh, scales)
        cd, dn = cosdist(normal0, n, F32(reg_weight * reg_weight_n / P))
        dqh = dqh + _get_normal_bwd(qh, saved, dn)
        reg_loss = F32(nrm + reg_weight_n * cd.mean(dtype=np.float64))
    # F.normalize backward
    draw = dqh / den
    dden = -(dqh * (qh / den)).sum(-1, keepdims=True, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        draw = draw + np.where((nq >= F32(1e-12)) & (nq > 0), raw["rotation"] * (dden / nq), F32(0))
    g["rotation"] = draw.astype(F32)
    lrs = {"xyz": lr["position_lr"], "features_dc": lr["feature_lr"], "features_rest": lr["feature_lr"] / 20.0,
           "opacity": lr["opacity_lr"], "scaling": lr["scaling_lr"], "rotation": lr["rotation_lr"]}
    new_raw, new_state = {}, {}
    for k in lrs:
        m, v = state.get(k, (np.zeros_like(raw[k]), np.zeros_like(raw[k])))
        new_raw[k], m, v = adam_update(raw[k], g[k], m, v, lrs[k], step)
        new_state[k] = (m, v)
    return new_raw, new_state, g, reg_loss
