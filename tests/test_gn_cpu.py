"""CPU: the numpy oracle of the fused dense-tracker Gauss-Newton step against goldens produced by the reference's own
optimizer.py (tests/golden/make_golden_gn.py)."""
import importlib.util
import os

import numpy as np
import pytest

from util import GOLDEN_DIR, rel_err
from oracle import gn_oracle as go

_spec = importlib.util.spec_from_file_location("make_golden_gn", os.path.join(GOLDEN_DIR, "make_golden_gn.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_matches_reference_golden(name):
    model, frame, intr, T, dx = mg.case_inputs(name)
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    coords, Jc = go.projective_transform(T, model["disp"], intr)
    ok = np.isfinite(gold["coords"]).all(-1)
    assert rel_err(coords[ok], gold["coords"][ok]) <= 1e-5
    A, b, n = go.icp_terms(model, frame, T, coords, mg.ANGLE_THRES, mg.DIST_THRES)
    assert n == int(gold["n_icp"]) and n > 500
    assert rel_err(A, gold["A_icp"]) <= 1e-4 and rel_err(b, gold["b_icp"]) <= 1e-4
    A, b, n = go.rgb_terms(model, frame, coords, Jc)
    assert n == int(gold["n_rgb"]) and n > 500
    assert rel_err(A, gold["A_rgb"]) <= 1e-4 and rel_err(b, gold["b_rgb"]) <= 1e-4
    assert rel_err(go.update_transform(T, dx), gold["T_updated"]) <= 1e-6
    assert rel_err(go.update_transform(T, dx * np.float32(1e-5)), gold["T_updated_small"]) <= 1e-6


def test_gn_steps_reduce_the_icp_residual():
    """A few oracle GN steps from a perturbed pose pull the point-to-plane residual down (sanity of signs / update)."""
    model, frame, intr, T, dx = mg.case_inputs("gn_96x72")
    frame = dict(frame)
    frame["vertex"] = np.nan_to_num(frame["vertex"], nan=2.0)
    cost = []
    for _ in range(4):
        coords, _ = go.projective_transform(T, model["disp"], intr)
        R, t = T[:3, :3], T[:3, 3]
        vprev = model["vertex"].reshape(-1, 3) @ R.T + t
        vcurr = go.sample_nearest(frame["vertex"], coords, True).reshape(-1, 3)
        ncurr = np.nan_to_num(go.sample_nearest(frame["normal"], coords, True).reshape(-1, 3))
        cost.append(float(np.abs((ncurr * (vcurr - vprev)).sum(-1)).mean()))
        T = go.gn_step(model, frame, intr, T, mg.ANGLE_THRES, mg.DIST_THRES, True, 1e-4, 1e-6, 0.01, 0.001)["T_new"]
    assert cost[-1] < cost[0]


def test_product_math_matches_oracle_on_cpu(hostemu):
    """The product's per-pixel rows (eggfusion_b200/csrc/egt_gn_math.cuh, compiled for the CPU) against the oracle and
    the reference goldens: same valid-pixel sets, J^T J / J^T r within 1e-4."""
    import ctypes
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from eggfusion_b200._lib import Level
    for name in mg.CASES:
        model, frame, intr, T, dx = mg.case_inputs(name)
        gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        H, W = model["mask"].shape[:2]
        keep = [np.ascontiguousarray(model["disp"], np.float32), np.ascontiguousarray(model["vertex"], np.float32),
                np.ascontiguousarray(model["normal"], np.float32), np.ascontiguousarray(model["mask"]).view(np.uint8),
                np.ascontiguousarray(model["intensity"], np.float32), np.ascontiguousarray(frame["vertex"], np.float32),
                np.ascontiguousarray(frame["normal"], np.float32), np.ascontiguousarray(frame["mask"]).view(np.uint8),
                np.ascontiguousarray(frame["intensity"], np.float32), np.ascontiguousarray(frame["grad"], np.float32)]
        lv = Level(W, H, float(intr[0]), float(intr[1]), float(intr[2]), float(intr[3]), *[a.ctypes.data for a in keep])
        sums = np.zeros(56, np.float64)
        Tc = np.ascontiguousarray(T, np.float32)
        hostemu.emu_gn_accumulate(ctypes.byref(lv), Tc.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(mg.ANGLE_THRES),
                                  ctypes.c_float(mg.DIST_THRES), ctypes.c_int(1), sums.ctypes.data_as(ctypes.c_void_p))

        def tri(v):
            M = np.zeros((6, 6))
            k = 0
            for r in range(6):
                for c in range(r, 6):
                    M[r, c] = M[c, r] = v[k]
                    k += 1
            return M
        assert int(round(sums[27])) == int(gold["n_icp"]) and int(round(sums[55])) == int(gold["n_rgb"])
        assert rel_err(tri(sums[:21]), gold["A_icp"]) <= 1e-4 and rel_err(sums[21:27], gold["b_icp"]) <= 1e-4
        assert rel_err(tri(sums[28:49]), gold["A_rgb"]) <= 1e-4 and rel_err(sums[49:55], gold["b_rgb"]) <= 1e-4
