"""CPU restatement of the 6x6 solve of the dense tracker -- TEST INFRASTRUCTURE ONLY (tests/ may import it; the product
never does).

The reference solves (A + lm I) x = b on the host with Eigen:
    /root/reference/src/utils/cuda/src/tracking.cu:929-950  solveBlock:
        Eigen::Map<Eigen::MatrixXf> A_eigen(A_cpu.data_ptr<float>(), rows, cols);      // COLUMN-major view of torch's buffer
        x = (A_eigen + lm * Identity).colPivHouseholderQr().solve(b_eigen);
Eigen is an un-vendored, un-pinned third-party dependency (`#include <Eigen/Dense>`, tracking.cu:5; absent from
/root/reference and from this image), so its algorithm is restated here from the published sources of Eigen 3.4.0
(Eigen/src/QR/ColPivHouseholderQR.h: computeInPlace / _solve_impl; Eigen/src/Householder/Householder.h:
makeHouseholderInPlace / applyHouseholderOnTheLeft), in the reference's precision (float32):

  * column pivoting on the largest remaining column norm, with LAPACK-style norm down-dating and recomputation when the
    down-dated value loses accuracy (norm_downdate_threshold = sqrt(eps));
  * the number of non-zero pivots is frozen the first time the largest remaining squared column norm falls below
    threshold_helper * (rows - k), threshold_helper = max column norm * eps / rows;
  * solve: c = Q^T b (the first `nonzero_pivots` reflectors), back-substitution on the leading nonzero_pivots x
    nonzero_pivots triangle, x[perm[i]] = c[i], the other components 0 (a basic least-squares solution for
    rank-deficient systems -- NOT zeros, NOT the minimum-norm solution).

Pin status: there is no Eigen here to run, and the reference holds no test vector for this function, so this oracle is
pinned only on properties (exact solutions of well-conditioned systems, residual optimality and the zero pattern on
rank-deficient ones, agreement with float64 within the float32 error bound): **parity unpinned** against Eigen's own
rounding (its vectorised reductions sum in a different order; differences are last-bit).
"""
import numpy as np


def _make_householder(x, dt):
    """Householder.h makeHouseholderInPlace: returns (essential, tau, beta) for vector x (x[0] = c0)."""
    c0 = x[0]
    tail = x[1:]
    tail_sq = dt(np.sum(tail * tail, dtype=dt))
    tol = np.finfo(dt).tiny
    if tail_sq <= tol:
        return np.zeros_like(tail), dt(0), c0
    beta = dt(np.sqrt(dt(c0 * c0) + tail_sq))
    if c0 >= 0:
        beta = -beta
    essential = (tail / dt(c0 - beta)).astype(dt)
    tau = dt(dt(beta - c0) / beta)
    return essential, tau, beta


def _apply_left(M, essential, tau, dt):
    """Householder.h applyHouseholderOnTheLeft on M (rows x cols), in place."""
    if M.shape[0] == 1:
        M *= dt(1) - tau
    elif tau != 0:
        bottom = M[1:, :]
        tmp = (essential @ bottom).astype(dt)
        tmp = (tmp + M[0, :]).astype(dt)
        M[0, :] -= (tau * tmp).astype(dt)
        bottom -= np.outer(essential, (tau * tmp).astype(dt)).astype(dt)


def colpiv_householder_qr_solve(A, b, lm=0.0, dtype=np.float32, column_major_buffer=True):
    """x = (A' + lm I).colPivHouseholderQr().solve(b) where A' is the n x n buffer `A` (as torch lays it out, row-major)
    read the way the reference reads it: column-major, i.e. A' = A^T (column_major_buffer=True)."""
    dt = dtype
    A = np.asarray(A, dtype=dt)
    n = A.shape[0]
    qr = (A.T.copy() if column_major_buffer else A.copy()).astype(dt)
    qr += (dt(lm) * np.eye(n, dtype=dt)).astype(dt)
    c = np.asarray(b, dtype=dt).reshape(-1).copy()
    rows = cols = size = n
    eps = np.finfo(dt).eps
    norms_upd = np.array([np.sqrt(np.sum(qr[:, j] * qr[:, j], dtype=dt)) for j in range(cols)], dtype=dt)
    norms_dir = norms_upd.copy()
    threshold_helper = dt(dt(norms_upd.max() * eps) / dt(rows))
    downdate_thr = dt(np.sqrt(eps))
    nonzero_pivots = size
    perm = list(range(cols))
    hcoef = np.zeros(size, dtype=dt)
    for k in range(size):
        big = int(np.argmax(norms_upd[k:])) + k
        big_sq = dt(norms_upd[big] * norms_upd[big])
        if nonzero_pivots == size and big_sq < dt(threshold_helper * dt(rows - k)):
            nonzero_pivots = k
        if big != k:
            qr[:, [k, big]] = qr[:, [big, k]]
            norms_upd[[k, big]] = norms_upd[[big, k]]
            norms_dir[[k, big]] = norms_dir[[big, k]]
            perm[k], perm[big] = perm[big], perm[k]
        essential, tau, beta = _make_householder(qr[k:, k].copy(), dt)
        qr[k + 1:, k] = essential
        qr[k, k] = beta
        hcoef[k] = tau
        if k + 1 < cols:
            sub = qr[k:, k + 1:]
            _apply_left(sub, essential, tau, dt)
        for j in range(k + 1, cols):
            if norms_upd[j] != 0:
                temp = dt(abs(qr[k, j]) / norms_upd[j])
                temp = dt(dt(1 + temp) * dt(1 - temp))
                temp = dt(0) if temp < 0 else temp
                ratio = dt(norms_upd[j] / norms_dir[j])
                temp2 = dt(temp * dt(ratio * ratio))
                if temp2 <= downdate_thr:
                    norms_dir[j] = dt(np.sqrt(np.sum(qr[k + 1:, j] * qr[k + 1:, j], dtype=dt)))
                    norms_upd[j] = norms_dir[j]
                else:
                    norms_upd[j] = dt(norms_upd[j] * dt(np.sqrt(temp)))
    x = np.zeros(n, dtype=dt)
    if nonzero_pivots == 0:
        return x, 0
    for k in range(nonzero_pivots):                      # c = Q^T b
        v = c[k:].reshape(-1, 1)
        _apply_left(v, qr[k + 1:, k].copy(), hcoef[k], dt)
    y = c[:nonzero_pivots].copy()
    for i in range(nonzero_pivots - 1, -1, -1):          # R y = c (upper triangular, leading block)
        s = y[i]
        for j in range(i + 1, nonzero_pivots):
            s = dt(s - dt(qr[i, j] * y[j]))
        y[i] = dt(s / qr[i, i])
    for i in range(nonzero_pivots):
        x[perm[i]] = y[i]
    return x, nonzero_pivots
