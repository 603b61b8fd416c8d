"""GPU parity of the fused dense-tracker Gauss-Newton step (include/eggtrack.h egt_gn_*) through the C ABI: against
the numpy oracle and the goldens produced by the reference's own optimizer.py."""
import os
import types

import numpy as np
import pytest
import torch

from util import GOLDEN_DIR, rel_err
from oracle import gn_oracle as go
from test_gn_cpu import mg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def pyramid(m, intr, levels=1):
    lists = {k: [] for k in ("vertex", "normal", "mask", "intensity", "grad", "disp")}
    intrs = []
    for l in range(levels):
        s = 2 ** l
        for k in lists:
            lists[k].append(t(m[k][::s, ::s]))
        intrs.append(torch.from_numpy(intr / s))
    return types.SimpleNamespace(vertex_pyramid=lists["vertex"], normal_pyramid=lists["normal"],
                                 mask_pyramid=lists["mask"], intensity_pyramid=lists["intensity"],
                                 grad_pyramid=lists["grad"], disp_pyramid=lists["disp"], intrinsic_pyramid=intrs)


def unpack(system, sums):
    s = system.cpu().numpy()
    A, b = s[:36].reshape(6, 6), s[36:]
    q = sums.cpu().numpy()

    def tri(v):
        M = np.zeros((6, 6))
        k = 0
        for r in range(6):
            for c in range(r, 6):
                M[r, c] = M[c, r] = v[k]
                k += 1
        return M
    return A, b, tri(q[:21]), q[21:27], int(round(q[27])), tri(q[28:49]), q[49:55], int(round(q[55]))


@pytest.mark.parametrize("name", list(mg.CASES))
def test_gn_step_matches_oracle_and_reference_golden(name):
    from eggfusion_b200 import tracking as TR
    model, frame, intr, T, dx = mg.case_inputs(name)
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = TR.TrackingConfig(angle_threshold=mg.ANGLE_THRES, distance_threshold=mg.DIST_THRES)
    trk = TR.DenseTracker(cfg, DEV)
    Tt = t(T).clone()
    dxd, conv = trk.tracking_optimization(pyramid(model, intr), pyramid(frame, intr), 0, Tt)
    A, b, A_icp, b_icp, n_icp, A_rgb, b_rgb, n_rgb = unpack(trk.system, trk.sums)
    assert n_icp == int(gold["n_icp"]) and n_rgb == int(gold["n_rgb"])
    assert rel_err(A_icp, gold["A_icp"]) <= 1e-4 and rel_err(b_icp, gold["b_icp"]) <= 1e-4
    assert rel_err(A_rgb, gold["A_rgb"]) <= 1e-4 and rel_err(b_rgb, gold["b_rgb"]) <= 1e-4
    o = go.gn_step(model, frame, intr, T, cfg.angle_threshold, cfg.distance_threshold, True, cfg.rgb_weight, cfg.lm,
                   cfg.residual_thres, cfg.dx_threshold)
    assert rel_err(A, o["A"]) <= 1e-5 and rel_err(b, o["b"]) <= 1e-5
    # the solve: the fused step eliminates in float64 (deliberately more accurate than the reference's fp32 Eigen QR);
    # it must agree with float64 on the SAME system, and with the restated Eigen algorithm (oracle/qr_oracle.py) within
    # that algorithm's own float32 error bound eps * cond(A)
    ref_dx = np.linalg.solve(A.astype(np.float64) + cfg.lm * np.eye(6), b.astype(np.float64))
    assert rel_err(dxd.cpu().numpy(), ref_dx) <= 1e-5
    from oracle.qr_oracle import colpiv_householder_qr_solve as qr_solve
    eig_dx, rank = qr_solve(A, b, cfg.lm)
    cond = np.linalg.cond(A.astype(np.float64) + cfg.lm * np.eye(6))
    assert rank == 6
    assert rel_err(dxd.cpu().numpy(), eig_dx) <= 20 * np.finfo(np.float32).eps * cond + 1e-6, (cond,)
    assert bool(conv) == o["converged"]
    assert rel_err(Tt.cpu().numpy(), go.update_transform(T, dxd.cpu().numpy())) <= 1e-6
    st = trk.status.cpu().numpy()
    assert st[2] == n_icp and st[3] == n_rgb


def test_update_transform_matches_reference_golden():
    """update_transform alone: feed a system whose solution is the golden's dx (A = I, b = dx, lm = 0)."""
    from eggfusion_b200 import _lib, tracking as TR
    lib = _lib.load()
    for name in mg.CASES:
        model, frame, intr, T, dx = mg.case_inputs(name)
        gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        for scale, key in ((1.0, "T_updated"), (1e-5, "T_updated_small")):
            sums = np.zeros(56)
            k = 0
            for r in range(6):
                for c in range(r, 6):
                    sums[k] = 1.0 if r == c else 0.0
                    k += 1
            sums[21:27] = dx * np.float32(scale)
            Tt = t(T).clone()
            _lib.check(lib.egt_gn_solve_update(t(sums).data_ptr(), 0.0, 0.0, 0.01, 0.001, Tt.data_ptr(), None, None, None,
                                               None))
            torch.cuda.synchronize()
            assert rel_err(Tt.cpu().numpy(), gold[key]) <= 1e-6, (name, key)


def test_track_loop_runs_without_sync_and_matches_oracle():
    from eggfusion_b200 import tracking as TR
    model, frame, intr, T, dx = mg.case_inputs("gn_96x72")
    cfg = TR.TrackingConfig(pyramid_level=2, pyramid_iters=(2, 2), angle_threshold=mg.ANGLE_THRES,
                            distance_threshold=mg.DIST_THRES)
    trk = TR.DenseTracker(cfg, DEV)
    pm, pf = pyramid(model, intr, 2), pyramid(frame, intr, 2)
    prev = t(np.eye(4, dtype=np.float32))
    curr, conv = trk.track(pm, pf, t(T), prev)
    # oracle: the same 4 steps
    To = T.copy()
    any_conv = False
    for l in range(2):
        level = 1 - l
        s = 2 ** level
        sub = lambda m: {k: np.ascontiguousarray(v[::s, ::s]) for k, v in m.items()}
        for _ in range(2):
            o = go.gn_step(sub(model), sub(frame), intr / s, To, cfg.angle_threshold, cfg.distance_threshold, True,
                           cfg.rgb_weight, cfg.lm, cfg.residual_thres, cfg.dx_threshold)
            To = o["T_new"]
            any_conv = any_conv or o["converged"]
    expect = To if any_conv else T
    assert bool(conv) == any_conv
    assert rel_err(curr.cpu().numpy(), expect) <= 1e-4
    with pytest.raises(RuntimeError, match="CUDA"):
        trk.tracking_optimization(pyramid(model, intr), pyramid(frame, intr), 0, torch.eye(4))
