// egt_qr.cuh -- column-pivoted Householder QR solve of one small dense system, fp32, host/device.
//
// What the reference's solveBlock does on the CPU through Eigen (/root/reference/src/utils/cuda/src/tracking.cu:929-950:
// `(A + lm I).colPivHouseholderQr().solve(b)` on a COLUMN-major view of torch's row-major buffer, i.e. on A^T), here on
// the device: same algorithm (Eigen 3.4 ColPivHouseholderQR::computeInPlace / _solve_impl: pivot on the largest remaining
// column norm with LAPACK-style norm down-dating, pivot count frozen by the threshold_helper rule, basic solution with
// zeros for the dropped pivots on rank-deficient systems), same precision.  oracle/qr_oracle.py restates the same
// published algorithm in numpy; tests/hostemu compiles this header for the CPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EGT_QR_HD __host__ __device__ inline
#else
#define EGT_QR_HD inline
#endif

#define EGT_QR_MAXN 16

// A: n x n buffer as torch lays it out (row-major); it is READ COLUMN-MAJOR like the reference reads it.
// Returns the number of non-zero pivots (the rank Eigen's solve uses).
EGT_QR_HD int egt_colpiv_qr_solve(const float* A, const float* b, float lm, float* x, int n) {
    float qr[EGT_QR_MAXN][EGT_QR_MAXN];   // qr[row][col] of M = A^T + lm I
    float c[EGT_QR_MAXN], nu[EGT_QR_MAXN], nd[EGT_QR_MAXN], hc[EGT_QR_MAXN];
    int perm[EGT_QR_MAXN];
    const float eps = 1.1920929e-07f;
    for (int r = 0; r < n; r++)
        for (int col = 0; col < n; col++) qr[r][col] = A[col * n + r] + (r == col ? lm : 0.f);
    for (int i = 0; i < n; i++) { c[i] = b[i]; perm[i] = i; x[i] = 0.f; hc[i] = 0.f; }
    float maxn = 0.f;
    for (int j = 0; j < n; j++) {
        float s = 0.f;
        for (int r = 0; r < n; r++) s += qr[r][j] * qr[r][j];
        nu[j] = nd[j] = sqrtf(s);
        maxn = nu[j] > maxn ? nu[j] : maxn;
    }
    const float threshold_helper = (maxn * eps) / (float)n;
    const float downdate_thr = sqrtf(eps);
    int nonzero = n;
    for (int k = 0; k < n; k++) {
        int big = k;
        for (int j = k + 1; j < n; j++)
            if (nu[j] > nu[big]) big = j;
        if (nonzero == n && nu[big] * nu[big] < threshold_helper * (float)(n - k)) nonzero = k;
        if (big != k) {
            for (int r = 0; r < n; r++) { const float t = qr[r][k]; qr[r][k] = qr[r][big]; qr[r][big] = t; }
            float t = nu[k]; nu[k] = nu[big]; nu[big] = t;
            t = nd[k]; nd[k] = nd[big]; nd[big] = t;
            const int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
        }
        // makeHouseholderInPlace on qr[k.., k]
        const float c0 = qr[k][k];
        float tail = 0.f;
        for (int r = k + 1; r < n; r++) tail += qr[r][k] * qr[r][k];
        float tau, beta;
        if (tail <= 1.17549435e-38f) {
            tau = 0.f;
            beta = c0;
            for (int r = k + 1; r < n; r++) qr[r][k] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tail);
            if (c0 >= 0.f) beta = -beta;
            const float d = c0 - beta;
            for (int r = k + 1; r < n; r++) qr[r][k] = qr[r][k] / d;
            tau = (beta - c0) / beta;
        }
        qr[k][k] = beta;
        hc[k] = tau;
        // applyHouseholderOnTheLeft to the remaining columns
        if (n - k == 1) {
            for (int j = k + 1; j < n; j++) qr[k][j] *= 1.f - tau;
        } else if (tau != 0.f) {
            for (int j = k + 1; j < n; j++) {
                float t = 0.f;
                for (int r = k + 1; r < n; r++) t += qr[r][k] * qr[r][j];
                t += qr[k][j];
                qr[k][j] -= tau * t;
                for (int r = k + 1; r < n; r++) qr[r][j] -= qr[r][k] * (tau * t);
            }
        }
        // down-date the remaining column norms
        for (int j = k + 1; j < n; j++) {
            if (nu[j] != 0.f) {
                float temp = fabsf(qr[k][j]) / nu[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                const float ratio = nu[j] / nd[j];
                const float temp2 = temp * (ratio * ratio);
                if (temp2 <= downdate_thr) {
                    float s = 0.f;
                    for (int r = k + 1; r < n; r++) s += qr[r][j] * qr[r][j];
                    nd[j] = sqrtf(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] *= sqrtf(temp);
                }
            }
        }
    }
    if (nonzero == 0) return 0;
    for (int k = 0; k < nonzero; k++) {   // c = Q^T b
        const float tau = hc[k];
        if (n - k == 1) {
            c[k] *= 1.f - tau;
        } else if (tau != 0.f) {
            float t = 0.f;
            for (int r = k + 1; r < n; r++) t += qr[r][k] * c[r];
            t += c[k];
            c[k] -= tau * t;
            for (int r = k + 1; r < n; r++) c[r] -= qr[r][k] * (tau * t);
        }
    }
    for (int i = nonzero - 1; i >= 0; i--) {   // back-substitution on the leading triangle
        float s = c[i];
        for (int j = i + 1; j < nonzero; j++) s -= qr[i][j] * c[j];
        c[i] = s / qr[i][i];
    }
    for (int i = 0; i < nonzero; i++) x[perm[i]] = c[i];
    return nonzero;
}
