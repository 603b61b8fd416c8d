// egt_gn.cu -- one Gauss-Newton step of the dense tracker, fused (include/eggtrack.h).
//
// Reference (all PyTorch there): Tracker.tracking_optimization (/root/reference/src/core/tracker.py:194-251) =
//   projective_transform (src/core/optimizer.py:131-180)  per-pixel warp + 2x6 Jacobian, materialised as [H,W,2,6]
//   icp_optimization     (optimizer.py:317-377)           point-to-plane residual, 2 nearest grid_samples, 6 masks,
//                                                         boolean-mask compaction (host sync), J^T J, J^T r
//   rgb_optimization     (optimizer.py:278-315)           photometric residual, 3 grid_samples, matmul Ji @ Jc, ...
//   solve_block          (src/utils/cuda/src/tracking.cu:929-950: GPU -> CPU Eigen QR -> GPU)
//   4 x .item() for the convergence test, update_transform (optimizer.py:426-441)
// ~70 launches and 7 host round trips per step, 9 steps per frame.  Here: k_gn_accumulate makes ONE pass over the
// level's pixels (every map read once, Jacobian rows live in registers, the 2 x (21 + 6) sums + 2 counts are reduced
// in registers -> shuffles -> one double atomic per CTA and value), k_gn_solve_update (one thread) combines the terms,
// solves the 6x6 system, evaluates the convergence test and applies update_transform to the pose ON THE DEVICE, so
// the whole pyramid loop runs without a host sync.  HBM-bound streaming: ~120 B per pixel.
#include "egs_common.cuh"
#include "../../include/eggtrack.h"

namespace {

#define GN_CTA 128
#define GN_SUMS 56   // icp: 21 (upper triangle of J^T J) + 6 (J^T r) + 1 (count); rgb: the same

struct Pose {
    float m[16];
};

// F.grid_sample(align_corners=True): [-1, 1] -> [0, size - 1]
__device__ __forceinline__ float unnormalize(float g, int size) { return (g + 1.f) * 0.5f * (float)(size - 1); }

// accumulate the upper triangle of J J^T (21), J r (6) and the count
__device__ __forceinline__ void accumulate(float (&acc)[28], const float (&J)[6], float r) {
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) { acc[k] = fmaf(J[a], J[b], acc[k]); k++; }
#pragma unroll
    for (int a = 0; a < 6; a++) acc[21 + a] = fmaf(J[a], r, acc[21 + a]);
    acc[27] += 1.f;
}

__global__ void __launch_bounds__(GN_CTA)
k_gn_accumulate(egt_level lv, const float* __restrict__ transform, float sine_thres, float dist_thres, int use_rgb,
                int px_per_thread, double* __restrict__ sums) {
    __shared__ float s_T[16];
    __shared__ float s_red[GN_CTA / 32][GN_SUMS];
    if (threadIdx.x < 16) s_T[threadIdx.x] = transform[threadIdx.x];
    __syncthreads();
    const float* T = s_T;
    const int W = lv.width, H = lv.height;
    const long long n = (long long)W * H;
    const float fx = lv.fx, fy = lv.fy, cx = lv.cx, cy = lv.cy;
    float icp[28], rgb[28];
#pragma unroll
    for (int k = 0; k < 28; k++) { icp[k] = 0.f; rgb[k] = 0.f; }

    const long long first = ((long long)blockIdx.x * px_per_thread) * GN_CTA + threadIdx.x;
    for (int it = 0; it < px_per_thread; it++) {
        const long long p = first + (long long)it * GN_CTA;
        if (p >= n) break;
        const int y = (int)(p / W), x = (int)(p - (long long)y * W);
        // ---- projective_transform (optimizer.py:131-180)
        const float us = ((float)x - cx) / fx, vs = ((float)y - cy) / fy, ds = lv.model_disp[p];
        float ut = T[0] * us + T[1] * vs + T[2] + T[3] * ds;
        float vt = T[4] * us + T[5] * vs + T[6] + T[7] * ds;
        const float zt = T[8] * us + T[9] * vs + T[10] + T[11] * ds;
        float dt = T[12] * us + T[13] * vs + T[14] + T[15] * ds;
        ut = ut / zt; vt = vt / zt; dt = dt / zt;
        const float gx = 2.f * (fx * ut + cx) / (float)(W - 1) - 1.f;
        const float gy = 2.f * (fy * vt + cy) / (float)(H - 1) - 1.f;
        // both terms need the warped pixel inside (the ICP bound 0.98 is the wider one); NaN coordinates fail it
        if (!(gx > -0.98f && gx < 0.98f && gy > -0.98f && gy < 0.98f)) continue;
        if (!lv.model_mask[p]) continue;                       // mask_prev gates both terms
        const float ix = unnormalize(gx, W), iy = unnormalize(gy, H);

        // ---- icp_optimization (optimizer.py:317-377)
        if (lv.frame_mask[p]) {                                 // mask_curr at the SAME pixel, not warped (as the reference)
            const float* vp = lv.model_vertex + 3 * p;
            const float* np_ = lv.model_normal + 3 * p;
            const float v0 = vp[0], v1 = vp[1], v2 = vp[2], n0 = np_[0], n1 = np_[1], n2 = np_[2];
            const float pv0 = T[0] * v0 + T[1] * v1 + T[2] * v2 + T[3];
            const float pv1 = T[4] * v0 + T[5] * v1 + T[6] * v2 + T[7];
            const float pv2 = T[8] * v0 + T[9] * v1 + T[10] * v2 + T[11];
            const float pn0 = T[0] * n0 + T[1] * n1 + T[2] * n2;
            const float pn1 = T[4] * n0 + T[5] * n1 + T[6] * n2;
            const float pn2 = T[8] * n0 + T[9] * n1 + T[10] * n2;
            // nearest, padding border, align_corners: clip then round half to even
            const int sx = (int)nearbyintf(fminf(fmaxf(ix, 0.f), (float)(W - 1)));
            const int sy = (int)nearbyintf(fminf(fmaxf(iy, 0.f), (float)(H - 1)));
            const long long q = (long long)sy * W + sx;
            const float* vc = lv.frame_vertex + 3 * q;
            const float* nc = lv.frame_normal + 3 * q;
            const float c0 = nc[0], c1 = nc[1], c2 = nc[2];
            const float d0 = vc[0] - pv0, d1 = vc[1] - pv1, d2 = vc[2] - pv2;
            const float x0 = c1 * pn2 - c2 * pn1, x1 = c2 * pn0 - c0 * pn2, x2 = c0 * pn1 - c1 * pn0;   // cross(ncurr, nprev)
            const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2), sine = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
            const bool nan_ok = x0 == x0 && x1 == x1 && x2 == x2;
            if (nan_ok && pv2 > 0.f && sine < sine_thres && dist < dist_thres) {
                const float r = c0 * d0 + c1 * d1 + c2 * d2;
                const float J[6] = {c0, c1, c2, pv1 * c2 - pv2 * c1, pv2 * c0 - pv0 * c2, pv0 * c1 - pv1 * c0};   // cross(vprev, ncurr)
                accumulate(icp, J, r);
            }
        }
        // ---- rgb_optimization (optimizer.py:278-315)
        if (use_rgb && gx > -0.90f && gx < 0.90f && gy > -0.90f && gy < 0.90f && lv.frame_grad[3 * p + 2] > 1.f) {
            // mask_curr: nearest, padding zeros
            const int mx = (int)nearbyintf(ix), my = (int)nearbyintf(iy);
            const bool mcur = mx >= 0 && mx < W && my >= 0 && my < H && lv.frame_mask[(long long)my * W + mx] != 0;
            if (mcur) {
                // bilinear, padding zeros, align_corners
                const float fx0 = floorf(ix), fy0 = floorf(iy);
                const int x0 = (int)fx0, y0 = (int)fy0;
                const float tx = ix - fx0, ty = iy - fy0;
                const float w00 = (1.f - tx) * (1.f - ty), w10 = tx * (1.f - ty), w01 = (1.f - tx) * ty, w11 = tx * ty;
                float sI = 0.f, sgx = 0.f, sgy = 0.f;
                auto tap = [&](int xx, int yy, float w) {
                    if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
                        const long long q = (long long)yy * W + xx;
                        sI = fmaf(lv.frame_intensity[q], w, sI);
                        sgx = fmaf(lv.frame_grad[3 * q], w, sgx);
                        sgy = fmaf(lv.frame_grad[3 * q + 1], w, sgy);
                    }
                };
                tap(x0, y0, w00); tap(x0 + 1, y0, w10); tap(x0, y0 + 1, w01); tap(x0 + 1, y0 + 1, w11);
                const float r = lv.model_intensity[p] - sI;
                // J = Ji (1x2) @ Jc (2x6), Jc rows as in projective_transform
                const float a = sgx * fx, b = sgy * fy;
                const float J[6] = {a * dt, b * dt, -(a * ut + b * vt) * dt, -a * ut * vt - b * (1.f + vt * vt),
                                    a * (1.f + ut * ut) + b * ut * vt, -a * vt + b * ut};
                accumulate(rgb, J, r);
            }
        }
    }
    // ---- CTA reduction: shuffles inside a warp, shared memory across the warps, one double atomic per value
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 28; k++) {
        float a = icp[k], b = rgb[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, d);
            b += __shfl_xor_sync(0xffffffffu, b, d);
        }
        if (lane == 0) { s_red[warp][k] = a; s_red[warp][28 + k] = b; }
    }
    __syncthreads();
    if (threadIdx.x < GN_SUMS) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < GN_CTA / 32; wv++) s += (double)s_red[wv][threadIdx.x];
        if (s != 0.0) atomicAdd(sums + threadIdx.x, s);
    }
}

// so3_to_SO3 (src/utils/camera_utils.py:18-28)
__device__ void so3_exp(const float w[3], float R[9]) {
    const float W[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
    float W2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) W2[3 * i + j] = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
    const float angle = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    float a, b;
    if (angle < 1e-5f) { a = 1.f; b = 0.5f; }
    else { a = sinf(angle) / angle; b = (1.f - cosf(angle)) / (angle * angle); }
    for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.f : 0.f) + a * W[i] + b * W2[i];
}

__global__ void k_gn_solve_update(const double* __restrict__ sums, float rgb_weight, float lm, float residual_thres,
                                  float dx_thres, float* __restrict__ transform, float* __restrict__ dx_out,
                                  float* __restrict__ A_out, int32_t* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // A = A_icp + rgb_weight * A_rgb, b likewise (tracker.py:229-230), formed in fp32 like the reference
    float A[6][6], b[6];
    int k = 0;
    for (int r = 0; r < 6; r++)
        for (int c = r; c < 6; c++, k++) {
            const float v = (float)sums[k] + rgb_weight * (float)sums[28 + k];
            A[r][c] = v; A[c][r] = v;
        }
    for (int r = 0; r < 6; r++) b[r] = (float)sums[21 + r] + rgb_weight * (float)sums[28 + 21 + r];
    if (A_out) {
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) A_out[6 * r + c] = A[r][c];
        for (int r = 0; r < 6; r++) A_out[36 + r] = b[r];
    }
    // (A + lm I) x = b, Gaussian elimination with partial pivoting in fp64 (as egt_solve_block)
    double M[6][7];
    for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) M[r][c] = (double)A[r][c] + (r == c ? (double)lm : 0.0);
        M[r][6] = (double)b[r];
    }
    bool singular = false;
    for (int kk = 0; kk < 6; kk++) {
        int p = kk;
        double best = fabs(M[kk][kk]);
        for (int r = kk + 1; r < 6; r++)
            if (fabs(M[r][kk]) > best) { best = fabs(M[r][kk]); p = r; }
        if (!(best > 0.0)) { singular = true; break; }
        if (p != kk)
            for (int c = kk; c <= 6; c++) { const double t = M[kk][c]; M[kk][c] = M[p][c]; M[p][c] = t; }
        for (int r = kk + 1; r < 6; r++) {
            const double f = M[r][kk] / M[kk][kk];
            for (int c = kk; c <= 6; c++) M[r][c] -= f * M[kk][c];
        }
    }
    float dx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (!singular) {
        double sol[6];
        for (int r = 5; r >= 0; r--) {
            double s = M[r][6];
            for (int c = r + 1; c < 6; c++) s -= M[r][c] * sol[c];
            sol[r] = s / M[r][r];
        }
        for (int r = 0; r < 6; r++) dx[r] = (float)sol[r];
    }
    // convergence test (tracker.py:235-250)
    const int icp_cnt = (int)(sums[27] + 0.5), rgb_cnt = (int)(sums[55] + 0.5);
    float bn = 0.f, dn = 0.f;
    for (int r = 0; r < 6; r++) { bn += b[r] * b[r]; dn += dx[r] * dx[r]; }
    bn = sqrtf(bn); dn = sqrtf(dn);
    const float residual_est = bn / fmaxf(1.f, sqrtf((float)(icp_cnt + rgb_cnt)));
    const int converged = (residual_est < residual_thres) && (dn < dx_thres);
    // update_transform (optimizer.py:426-441): R <- exp(dx[3:]) R, t <- dx[:3] + t  (t is NOT rotated: as the reference)
    float dR[9];
    so3_exp(dx + 3, dR);
    float Rn[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Rn[3 * i + j] = dR[3 * i] * transform[j] + dR[3 * i + 1] * transform[4 + j] + dR[3 * i + 2] * transform[8 + j];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) transform[4 * i + j] = Rn[3 * i + j];
        transform[4 * i + 3] = dx[i] + transform[4 * i + 3];
    }
    if (dx_out)
        for (int r = 0; r < 6; r++) dx_out[r] = dx[r];
    if (status) {
        status[0] |= converged;     // dense_converged: any step of the frame (tracker.py:163-164)
        status[1] = converged;
        status[2] = icp_cnt;
        status[3] = rgb_cnt;
    }
}

#define EGT_TRY(expr)                              \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)
} // namespace

extern "C" {

EGS_API int egt_gn_accumulate(const egt_level* lv, const float* transform, float angle_thres_deg, float dist_thres,
                              int32_t use_rgb, double* sums, void* stream) {
    if (!lv || !transform || !sums || lv->width <= 1 || lv->height <= 1) return EGS_E_BADARG;
    if (!lv->model_disp || !lv->model_vertex || !lv->model_normal || !lv->model_mask || !lv->frame_vertex ||
        !lv->frame_normal || !lv->frame_mask)
        return EGS_E_BADARG;
    if (use_rgb && (!lv->model_intensity || !lv->frame_intensity || !lv->frame_grad)) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    EGT_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * EGT_GN_SUMS, s));
    const long long n = (long long)lv->width * lv->height;
    // enough CTAs to fill 148 SMs several times over, at most 8 pixels per thread (amortises the 56-value reduction)
    int ppt = 8;
    while (ppt > 1 && (n + (long long)GN_CTA * ppt - 1) / ((long long)GN_CTA * ppt) < 148 * 4) ppt >>= 1;
    const long long ctas = (n + (long long)GN_CTA * ppt - 1) / ((long long)GN_CTA * ppt);
    const float sine_thres = (float)((double)angle_thres_deg * 3.14159265358979323846 / 180.0);
    k_gn_accumulate<<<(unsigned)ctas, GN_CTA, 0, s>>>(*lv, transform, sine_thres, dist_thres, use_rgb, ppt, sums);
    EGT_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egt_gn_solve_update(const double* sums, float rgb_weight, float lm, float residual_thres, float dx_thres,
                                float* transform, float* dx_out, float* system_out, int32_t* status, void* stream) {
    if (!sums || !transform) return EGS_E_BADARG;
    k_gn_solve_update<<<1, 32, 0, (cudaStream_t)stream>>>(sums, rgb_weight, lm, residual_thres, dx_thres, transform,
                                                          dx_out, system_out, status);
    EGT_TRY(cudaGetLastError());
    return 0;
}

} // extern "C"
