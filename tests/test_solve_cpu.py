"""The 6x6 solve of the dense tracker, CPU side: the oracle restatement of Eigen's colPivHouseholderQr
(oracle/qr_oracle.py; the reference: src/utils/cuda/src/tracking.cu:929-950) on known answers, and the PRODUCT's
host/device implementation (eggfusion_b200/csrc/egt_qr.cuh, compiled for the CPU by tests/hostemu) against the oracle."""
import ctypes

import numpy as np
import pytest

from oracle.qr_oracle import colpiv_householder_qr_solve as qr_solve


def _systems():
    rng = np.random.default_rng(11)
    out = []
    for i in range(6):   # Gauss-Newton normal matrices J^T J of growing condition number
        J = rng.normal(size=(60, 6)) * np.array([1, 1, 1, 10.0 ** i, 1, 10.0 ** (-i / 2)])
        out.append(("spd_cond%d" % i, (J.T @ J).astype(np.float32), rng.normal(size=6).astype(np.float32), 1e-6))
    out.append(("nonsymmetric", rng.normal(size=(6, 6)).astype(np.float32), rng.normal(size=6).astype(np.float32), 0.0))
    B = rng.normal(size=(4, 6))
    A = (B.T @ B).astype(np.float32)            # rank 4
    out.append(("rank4", A, (A @ rng.normal(size=6)).astype(np.float32), 0.0))
    out.append(("zero_plus_lm", np.zeros((6, 6), np.float32), rng.normal(size=6).astype(np.float32), 1e-6))
    out.append(("n3", rng.normal(size=(3, 3)).astype(np.float32), rng.normal(size=3).astype(np.float32), 1e-6))
    return out


@pytest.mark.parametrize("name,A,b,lm", _systems(), ids=[s[0] for s in _systems()])
def test_oracle_known_answers(name, A, b, lm):
    x, rank = qr_solve(A, b, lm)
    n = A.shape[0]
    M = A.T.astype(np.float64) + lm * np.eye(n)          # the column-major view the reference solves
    if name == "zero_plus_lm":                          # the reference always damps: (0 + lm I) x = b
        assert rank == n and np.allclose(x, b / np.float32(1e-6), rtol=1e-6)
        return
    if name == "rank4":
        assert rank == 4 and int((x == 0).sum()) == 2    # Eigen's basic solution: dropped pivots stay zero
        assert np.abs(M @ x - b).max() <= 1e-5 * np.abs(b).max()
        return
    cond = np.linalg.cond(M)
    if cond > 1e7:
        # numerically singular in float32: Eigen's pivot-count rule drops pivots and returns a basic solution
        assert rank < n and int((x == 0).sum()) == n - rank
        return
    assert rank == n
    x64 = np.linalg.solve(M, b.astype(np.float64))
    assert np.abs(x - x64).max() <= 50 * np.finfo(np.float32).eps * cond * np.abs(x64).max() + 1e-7


@pytest.mark.parametrize("name,A,b,lm", _systems(), ids=[s[0] for s in _systems()])
def test_product_qr_matches_oracle(hostemu, name, A, b, lm):
    n = A.shape[0]
    x = np.zeros(n, np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    hostemu.emu_colpiv_qr_solve.restype = ctypes.c_int
    rank = hostemu.emu_colpiv_qr_solve(np.ascontiguousarray(A).ctypes.data_as(fp), b.ctypes.data_as(fp), ctypes.c_float(lm),
                                       x.ctypes.data_as(fp), n)
    xo, ro = qr_solve(A, b, lm)
    assert rank == ro
    assert np.array_equal(x == 0, xo == 0)
    M = A.T.astype(np.float64) + lm * np.eye(n)
    cond = min(np.linalg.cond(M), 1e7) if rank == n else 1e4
    # same algorithm, same precision: they differ only by the summation order of the inner products
    assert np.abs(x - xo).max() <= 20 * np.finfo(np.float32).eps * cond * max(np.abs(xo).max(), 1e-30) + 1e-7
