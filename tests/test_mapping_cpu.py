"""CPU checks of the mapping-iteration glue: the numpy oracle against golden vectors produced by the
reference's own python code (tests/golden/make_golden_mapping.py), and the product's __host__ __device__ arithmetic
(eggfusion_b200/csrc/egm_math.cuh, compiled for the CPU in tests/hostemu) against the oracle."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest

from util import GOLDEN_DIR, rel_err
from oracle import mapping_oracle as mo

_spec = importlib.util.spec_from_file_location("make_golden_mapping", os.path.join(GOLDEN_DIR, "make_golden_mapping.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

TOL = 1e-4          # north_star: within 1e-4 rel fp32
NAMES = ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")


def run_oracle(name, gold=None):
    """The oracle over the case's iterations -> dict with the golden's keys.  With `gold`, every iteration starts from
    the reference's parameters and Adam state of the previous one (each step is checked on its own: one ulp in a
    parameter is 1e-4 of the (xyz - pos0) difference the regulariser gradient is made of)."""
    raw, its, (cw, dw, nw, rw, rwn), lr = mg.case_inputs(name)
    raw = {k: v.copy() for k, v in raw.items()}
    pos0 = raw["xyz"].copy()
    _, _, _, normal0 = mo.activate(raw["opacity"], raw["scaling"], raw["rotation"])
    out, state = {"normal0": normal0}, {}
    for k, it in enumerate(its):
        if gold is not None and k > 0:
            raw = {n: gold[f"param_{n}_{k - 1}"] for n in NAMES}
            state = {n: (gold[f"m_{n}_{k - 1}"], gold[f"v_{n}_{k - 1}"]) for n in NAMES}
        ls = mo.loss_seed(it["est_color"], it["est_depth"], it["est_normal"], it["ref_color"], it["ref_depth"],
                          it["ref_normal"], it["rgb_mask"], it["geo_mask"], cw, dw, nw)
        raw, state, g, reg = mo.adam_step(raw, it, state, lr, k + 1, rw, rwn, pos0, normal0)
        out[f"loss_{k}"] = np.float32(ls["image_loss"] + rw * reg)
        out[f"seed_color_{k}"], out[f"seed_depth_{k}"], out[f"seed_normal_{k}"] = ls["dL_dcolor"], ls["dL_ddepth"], ls["dL_dnormal"]
        for n in NAMES:
            out[f"grad_{n}_{k}"] = g[n]
            out[f"param_{n}_{k}"] = raw[n]
            out[f"m_{n}_{k}"], out[f"v_{n}_{k}"] = state[n]
        op, sc, rot, nrm = mo.activate(raw["opacity"], raw["scaling"], raw["rotation"])
        out[f"act_opacity_{k}"], out[f"act_scales_{k}"], out[f"act_rotations_{k}"], out[f"act_normal_{k}"] = op, sc, rot, nrm
    return out


def knife_edge_rows(name, gold):
    """Per iteration: surfels whose regulariser cosine sits within a few ulp of the clamp bound 1 - 1e-6.  There the
    reference's own fp32 rounding decides whether `clamp` passes a gradient (the anchors ARE the initial normals, so
    every cosine starts at ~1): a discontinuity of the reference, not comparable across implementations."""
    raw, its, (cw, dw, nw, rw, rwn), lr = mg.case_inputs(name)
    out = []
    for k in range(len(its)):
        r = raw if k == 0 else {n: gold[f"param_{n}_{k - 1}"] for n in NAMES}
        _, _, _, nrm = mo.activate(r["opacity"], r["scaling"], r["rotation"])
        c = np.abs((nrm.astype(np.float64) * gold["normal0"]).sum(1))
        out.append((np.abs((1 - c) - 1e-6) < 4e-7) if rw > 0 else np.zeros(c.shape, bool))
    return out


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_matches_reference_golden(name):
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    got = run_oracle(name, gold)
    assert set(gold.files) == set(got)
    raw, its, _w, _lr = mg.case_inputs(name)
    edge = knife_edge_rows(name, gold)
    worst = {}
    for k in gold.files:
        if k.startswith("param_"):
            # parameters: compare the UPDATE (param - start), otherwise 1e-4 of |param| would hide a wrong step
            continue
        a, b = got[k], gold[k]
        if "rotation" in k and k[-1].isdigit() and edge[int(k[-1])].any():
            a, b = a[~edge[int(k[-1])]], b[~edge[int(k[-1])]]
        e = rel_err(a, b)
        worst[k] = e
        assert e <= TOL, (k, e)
    for n in NAMES:
        prev = raw[n]
        for k in range(len(its)):
            keep = ~edge[k] if n == "rotation" else slice(None)
            dg, do = (gold[f"param_{n}_{k}"] - prev)[keep], (got[f"param_{n}_{k}"] - prev)[keep]
            if dg.size and np.abs(dg).max() > 0:
                assert rel_err(do, dg) <= 1e-3, (n, k, rel_err(do, dg))   # one ulp of the parameter vs the step size
            assert rel_err(got[f"param_{n}_{k}"][keep], gold[f"param_{n}_{k}"][keep]) <= 1e-6
            prev = gold[f"param_{n}_{k}"]


def test_seed_quirks_are_reproduced():
    """Properties of torch's autograd the product must keep: zero seed on exact ties (sign(0) = 0), no normal gradient
    where the cosine is clamped, 1/eps-scaled gradient on zero rendered normals, zero seeds outside the mask."""
    raw, its, (cw, dw, nw, rw, rwn), lr = mg.case_inputs("mapping_deg3")
    it = its[0]
    ls = mo.loss_seed(it["est_color"], it["est_depth"], it["est_normal"], it["ref_color"], it["ref_depth"],
                      it["ref_normal"], it["rgb_mask"], it["geo_mask"], cw, dw, nw)
    m = it["rgb_mask"] & it["geo_mask"]
    assert np.all(ls["dL_dcolor"][:, 4, :8] == 0)
    assert np.all(ls["dL_dnormal"][:, 6, :8] == 0)
    assert np.all(ls["dL_dcolor"][:, ~m] == 0) and np.all(ls["dL_dnormal"][:, ~m] == 0) and np.all(ls["dL_ddepth"][:, ~m] == 0)
    top = ls["dL_dnormal"][:, :3, :][:, m[:3, :]]
    assert np.abs(top).max() > 1e3      # r_hat / 1e-8 * normal_weight / count


def test_empty_mask_gives_nan_color_loss_and_zero_seeds():
    raw, its, (cw, dw, nw, rw, rwn), lr = mg.case_inputs("mapping_deg0_noreg")
    it = its[0]
    z = np.zeros_like(it["rgb_mask"])
    ls = mo.loss_seed(it["est_color"], it["est_depth"], it["est_normal"], it["ref_color"], it["ref_depth"],
                      it["ref_normal"], z, it["geo_mask"], cw, dw, nw)
    assert np.isnan(ls["color_loss"]) and ls["depth_loss"] == 0 and ls["normal_loss"] == 0
    assert not ls["dL_dcolor"].any() and not ls["dL_ddepth"].any() and not ls["dL_dnormal"].any()


# ---- the product's arithmetic (egm_math.cuh) on the CPU ------------------------------------------------------------
def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_product_math_matches_oracle_on_cpu(hostemu):
    raw, its, (cw, dw, nw, rw, rwn), lr = mg.case_inputs("mapping_deg3")
    it = its[0]
    # per-pixel normal term
    H, W = it["rgb_mask"].shape
    rn = np.ascontiguousarray(it["ref_normal"].reshape(-1, 3))
    en = np.ascontiguousarray(it["est_normal"].transpose(1, 2, 0).reshape(-1, 3))
    n = rn.shape[0]
    val, grad = np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    hostemu.emu_cosdist(ctypes.c_int(n), _fp(rn), _fp(en), ctypes.c_float(0.37), _fp(val), _fp(grad))
    cd, e = mo.cosdist(rn, en, np.float32(0.37))
    assert rel_err(val, cd) <= 1e-6 and rel_err(grad, e) <= 1e-5
    # per-surfel chain: activations, get_normal, regulariser + normalize backward, Adam
    P = raw["xyz"].shape[0]
    pos0 = raw["xyz"] + 1e-3 * np.random.default_rng(0).standard_normal((P, 3)).astype(np.float32)
    _, _, _, normal0 = mo.activate(raw["opacity"], raw["scaling"], raw["rotation"])
    normal0 = np.ascontiguousarray(normal0[::-1])       # mismatched anchors -> non-trivial regulariser gradient
    state = {k: (1e-4 * np.abs(raw[k]) + 1e-6, 1e-8 * raw[k] ** 2 + 1e-12) for k in NAMES}
    state = {k: (m.astype(np.float32), v.astype(np.float32)) for k, (m, v) in state.items()}
    new_raw, new_state, g, reg = mo.adam_step(raw, it, state, lr, 3, rw, rwn, pos0, normal0)
    nrm2 = np.float64(((pos0 - raw["xyz"]).astype(np.float64) ** 2).sum())
    geo = {k: raw[k].copy() for k in ("xyz", "opacity", "scaling", "rotation")}
    ms = {k: state[k][0].copy() for k in geo}
    vs = {k: state[k][1].copy() for k in geo}
    graw = {k: np.zeros_like(raw[k]) for k in geo}
    act_o, act_s, act_r = np.zeros((P, 1), np.float32), np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32)
    lrs = (ctypes.c_float * 4)(lr["position_lr"], lr["opacity_lr"], lr["scaling_lr"], lr["rotation_lr"])
    hostemu.emu_adam_geom(ctypes.c_int(P), ctypes.c_int(3), lrs, ctypes.c_float(rw), ctypes.c_float(rwn),
                          ctypes.c_double(nrm2), _fp(geo["xyz"]), _fp(geo["opacity"]), _fp(geo["scaling"]),
                          _fp(geo["rotation"]), _fp(it["G_xyz"]), _fp(it["G_opacity"]), _fp(it["G_scales"]),
                          _fp(it["G_rot"]), _fp(ms["xyz"]), _fp(vs["xyz"]), _fp(ms["opacity"]), _fp(vs["opacity"]),
                          _fp(ms["scaling"]), _fp(vs["scaling"]), _fp(ms["rotation"]), _fp(vs["rotation"]), _fp(pos0),
                          _fp(normal0), _fp(graw["xyz"]), _fp(graw["opacity"]), _fp(graw["scaling"]),
                          _fp(graw["rotation"]), _fp(act_o), _fp(act_s), _fp(act_r))
    for k in geo:
        assert rel_err(graw[k], g[k]) <= 1e-5, (k, rel_err(graw[k], g[k]))
        assert rel_err(geo[k] - raw[k], new_raw[k] - raw[k]) <= 1e-3, k
        assert rel_err(ms[k], new_state[k][0]) <= 1e-6 and rel_err(vs[k], new_state[k][1]) <= 1e-6
    o, s, r, _ = mo.activate(new_raw["opacity"], new_raw["scaling"], new_raw["rotation"])
    assert rel_err(act_o, o) <= 1e-6 and rel_err(act_s, s) <= 1e-6 and rel_err(act_r, r) <= 1e-6
