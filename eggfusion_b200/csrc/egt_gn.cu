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
// level's pixels (every map read once, Jacobian rows live in registers; the 2 x (21 + 6) sums + 2 counts are reduced
// by a transposing warp reduction -> one running scalar per lane -> shared memory -> one double atomic per CTA and
// value), k_gn_solve_update (one thread) combines the terms, solves the 6x6 system, evaluates the convergence test and
// applies update_transform to the pose ON THE DEVICE; egt_track_pyramid enqueues the whole coarse-to-fine loop
// (2 launches per step) without a host sync.  HBM-bound streaming: ~120 B per pixel.
#include "egt_gn_math.cuh"

namespace {

#define GN_CTA 128
#define GN_SUMS 56   // icp: 21 (upper triangle of J^T J) + 6 (J^T r) + 1 (count); rgb: the same

// The 28 sums of one term: upper triangle of J J^T (21), J r (6), count (1) -- padded to 32.
__device__ __forceinline__ void outer_terms(float (&v)[32], bool valid, const float (&J)[6], float r) {
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) { v[k] = valid ? J[a] * J[b] : 0.f; k++; }
#pragma unroll
    for (int a = 0; a < 6; a++) v[21 + a] = valid ? J[a] * r : 0.f;
    v[27] = valid ? 1.f : 0.f;
    v[28] = v[29] = v[30] = v[31] = 0.f;
}

// Transposing warp reduction: every lane brings 32 values; afterwards lane l holds the warp-wide sum of value l.
// 16 + 8 + 4 + 2 + 1 = 31 shuffles (a plain per-value butterfly would need 32 x 5).  Keeping 2 x 28 running sums per
// thread instead cost 120 registers (25 % occupancy) and left the level-0 pass latency-bound (65 us for 98 MB).
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
    const unsigned full = 0xffffffffu;
    float w16[16], w8[8], w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
    for (int i = 0; i < 16; i++) w16[i] = (b4 ? v[i + 16] : v[i]) + __shfl_xor_sync(full, b4 ? v[i] : v[i + 16], 16);
#pragma unroll
    for (int i = 0; i < 8; i++) w8[i] = (b3 ? w16[i + 8] : w16[i]) + __shfl_xor_sync(full, b3 ? w16[i] : w16[i + 8], 8);
#pragma unroll
    for (int i = 0; i < 4; i++) w4[i] = (b2 ? w8[i + 4] : w8[i]) + __shfl_xor_sync(full, b2 ? w8[i] : w8[i + 4], 4);
#pragma unroll
    for (int i = 0; i < 2; i++) w2[i] = (b1 ? w4[i + 2] : w4[i]) + __shfl_xor_sync(full, b1 ? w4[i] : w4[i + 2], 2);
    return (b0 ? w2[1] : w2[0]) + __shfl_xor_sync(full, b0 ? w2[0] : w2[1], 1);
}

__global__ void __launch_bounds__(GN_CTA)
k_gn_accumulate(egt_level lv, const float* __restrict__ transform, float sine_thres, float dist_thres, int use_rgb,
                int px_per_thread, double* __restrict__ sums) {
    __shared__ float s_T[16];
    __shared__ float s_red[GN_CTA / 32][64];
    if (threadIdx.x < 16) s_T[threadIdx.x] = transform[threadIdx.x];
    __syncthreads();
    const float* T = s_T;
    const long long n = (long long)lv.width * lv.height;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc_icp = 0.f, acc_rgb = 0.f;   // lane l: running sum of value l of each term

    const long long first = ((long long)blockIdx.x * px_per_thread) * GN_CTA + threadIdx.x;
    for (int it = 0; it < px_per_thread; it++) {
        const long long p = first + (long long)it * GN_CTA;
        if (p - lane >= n) break;   // warp-uniform: the whole warp is past the image
        const bool in_img = p < n;
        const long long pc = in_img ? p : 0;
        GnWarp w;
        egt_gn_warp(lv, T, pc, in_img, w);                     // projective_transform (optimizer.py:131-180)
        if (!__any_sync(0xffffffffu, w.base)) continue;
        float v[32], J[6], r;
        {   // icp_optimization (optimizer.py:317-377)
            const bool valid = egt_gn_icp_row(lv, T, pc, w, sine_thres, dist_thres, J, r);
            if (__any_sync(0xffffffffu, valid)) {
                outer_terms(v, valid, J, r);
                acc_icp += warp_transpose_sum32(v, lane);
            }
        }
        if (use_rgb) {   // rgb_optimization (optimizer.py:278-315)
            const bool valid = egt_gn_rgb_row(lv, pc, w, J, r);
            if (__any_sync(0xffffffffu, valid)) {
                outer_terms(v, valid, J, r);
                acc_rgb += warp_transpose_sum32(v, lane);
            }
        }
    }
    // ---- CTA combine: lane l of every warp holds value l of both terms; one double atomic per CTA and value
    s_red[warp][lane] = acc_icp;
    s_red[warp][32 + lane] = acc_rgb;
    __syncthreads();
    if (threadIdx.x < GN_SUMS) {
        const int src = threadIdx.x < 28 ? threadIdx.x : 32 + (threadIdx.x - 28);
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < GN_CTA / 32; wv++) s += (double)s_red[wv][src];
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

// One thread; every loop is fully unrolled with compile-time indices so the 6x7 fp64 system lives in registers (the
// rolled version indexed it dynamically -> local memory, 7.7 us of dependent L1 round trips per step).
__global__ void k_gn_solve_update(double* __restrict__ sums, int rezero, float rgb_weight, float lm,
                                  float residual_thres, float dx_thres, float* __restrict__ transform,
                                  float* __restrict__ dx_out, float* __restrict__ A_out, int32_t* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // A = A_icp + rgb_weight * A_rgb, b likewise (tracker.py:229-230), formed in fp32 like the reference
    double q[GN_SUMS];
#pragma unroll
    for (int i = 0; i < GN_SUMS; i++) q[i] = sums[i];
    if (rezero) {   // the next egt_gn_accumulate of a fused pyramid loop starts from zeroed sums without a memset
#pragma unroll
        for (int i = 0; i < GN_SUMS; i++) sums[i] = 0.0;
    }
    float A[6][6], b[6];
    {
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = r; c < 6; c++) {
                const float v = (float)q[k] + rgb_weight * (float)q[28 + k];
                A[r][c] = v; A[c][r] = v;
                k++;
            }
#pragma unroll
        for (int r = 0; r < 6; r++) b[r] = (float)q[21 + r] + rgb_weight * (float)q[28 + 21 + r];
    }
    if (A_out) {
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) A_out[6 * r + c] = A[r][c];
#pragma unroll
        for (int r = 0; r < 6; r++) A_out[36 + r] = b[r];
    }
    // (A + lm I) x = b, Gaussian elimination with partial pivoting in fp64 (as egt_solve_block)
    double M[6][7];
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) M[r][c] = (double)A[r][c] + (r == c ? (double)lm : 0.0);
        M[r][6] = (double)b[r];
    }
    bool singular = false;
#pragma unroll
    for (int kk = 0; kk < 6; kk++) {
        // bring the row with the largest |M[r][kk]|, r >= kk, to position kk by conditional swaps
#pragma unroll
        for (int r = kk + 1; r < 6; r++) {
            if (fabs(M[r][kk]) > fabs(M[kk][kk])) {
#pragma unroll
                for (int c = 0; c < 7; c++) { const double t = M[kk][c]; M[kk][c] = M[r][c]; M[r][c] = t; }
            }
        }
        if (!(fabs(M[kk][kk]) > 0.0)) singular = true;
        const double inv = 1.0 / M[kk][kk];
#pragma unroll
        for (int r = kk + 1; r < 6; r++) {
            const double f = M[r][kk] * inv;
#pragma unroll
            for (int c = kk; c < 7; c++) M[r][c] -= f * M[kk][c];
        }
    }
    float dx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (!singular) {
        double sol[6];
#pragma unroll
        for (int r = 5; r >= 0; r--) {
            double sacc = M[r][6];
#pragma unroll
            for (int c = r + 1; c < 6; c++) sacc -= M[r][c] * sol[c];
            sol[r] = sacc / M[r][r];
        }
#pragma unroll
        for (int r = 0; r < 6; r++) dx[r] = (float)sol[r];
    }
    // convergence test (tracker.py:235-250)
    const int icp_cnt = (int)(q[27] + 0.5), rgb_cnt = (int)(q[55] + 0.5);
    float bn = 0.f, dn = 0.f;
#pragma unroll
    for (int r = 0; r < 6; r++) { bn += b[r] * b[r]; dn += dx[r] * dx[r]; }
    bn = sqrtf(bn); dn = sqrtf(dn);
    const float residual_est = bn / fmaxf(1.f, sqrtf((float)(icp_cnt + rgb_cnt)));
    const int converged = (residual_est < residual_thres) && (dn < dx_thres);
    // update_transform (optimizer.py:426-441): R <- exp(dx[3:]) R, t <- dx[:3] + t  (t is NOT rotated: as the reference)
    float dR[9];
    so3_exp(dx + 3, dR);
    float Told[12];
#pragma unroll
    for (int i = 0; i < 12; i++) Told[i] = transform[i];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++)
            transform[4 * i + j] = dR[3 * i] * Told[j] + dR[3 * i + 1] * Told[4 + j] + dR[3 * i + 2] * Told[8 + j];
        transform[4 * i + 3] = dx[i] + Told[4 * i + 3];
    }
    if (dx_out) {
#pragma unroll
        for (int r = 0; r < 6; r++) dx_out[r] = dx[r];
    }
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

static int gn_accumulate(const egt_level* lv, const float* transform, float angle_thres_deg, float dist_thres,
                         int32_t use_rgb, double* sums, bool zero_first, cudaStream_t s) {
    if (!lv || !transform || !sums || lv->width <= 1 || lv->height <= 1) return EGS_E_BADARG;
    if (!lv->model_disp || !lv->model_vertex || !lv->model_normal || !lv->model_mask || !lv->frame_vertex ||
        !lv->frame_normal || !lv->frame_mask)
        return EGS_E_BADARG;
    if (use_rgb && (!lv->model_intensity || !lv->frame_intensity || !lv->frame_grad)) return EGS_E_BADARG;
    if (zero_first) EGT_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * EGT_GN_SUMS, s));
    const long long n = (long long)lv->width * lv->height;
    // enough CTAs to fill 148 SMs several times over, at most 8 pixels per thread
    int ppt = 8;
    while (ppt > 1 && (n + (long long)GN_CTA * ppt - 1) / ((long long)GN_CTA * ppt) < 148 * 8) ppt >>= 1;
    const long long ctas = (n + (long long)GN_CTA * ppt - 1) / ((long long)GN_CTA * ppt);
    const float sine_thres = (float)((double)angle_thres_deg * 3.14159265358979323846 / 180.0);
    k_gn_accumulate<<<(unsigned)ctas, GN_CTA, 0, s>>>(*lv, transform, sine_thres, dist_thres, use_rgb, ppt, sums);
    EGT_TRY(cudaGetLastError());
    return 0;
}

extern "C" {

EGS_API int egt_gn_accumulate(const egt_level* lv, const float* transform, float angle_thres_deg, float dist_thres,
                              int32_t use_rgb, double* sums, void* stream) {
    return gn_accumulate(lv, transform, angle_thres_deg, dist_thres, use_rgb, sums, true, (cudaStream_t)stream);
}

EGS_API int egt_gn_solve_update(const double* sums, float rgb_weight, float lm, float residual_thres, float dx_thres,
                                float* transform, float* dx_out, float* system_out, int32_t* status, void* stream) {
    if (!sums || !transform) return EGS_E_BADARG;
    k_gn_solve_update<<<1, 32, 0, (cudaStream_t)stream>>>(const_cast<double*>(sums), 0, rgb_weight, lm, residual_thres,
                                                          dx_thres, transform, dx_out, system_out, status);
    EGT_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egt_track_pyramid(const egt_level* levels, int32_t nlevel, const int32_t* iters, float angle_thres_deg,
                              float dist_thres, int32_t use_rgb, float rgb_weight, float lm, float residual_thres,
                              float dx_thres, float* transform, double* sums, float* dx_out, float* system_out,
                              int32_t* status, void* stream) {
    if (!levels || !iters || nlevel <= 0 || !transform || !sums || !status) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    EGT_TRY(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), s));
    EGT_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * EGT_GN_SUMS, s));
    for (int l = 0; l < nlevel; l++) {                       // tracker.py:157-159: coarse to fine
        const egt_level* lv = levels + (nlevel - 1 - l);
        for (int it = 0; it < iters[l]; it++) {
            const int rc = gn_accumulate(lv, transform, angle_thres_deg, dist_thres, use_rgb, sums, false, s);
            if (rc) return rc;
            k_gn_solve_update<<<1, 32, 0, s>>>(sums, 1, rgb_weight, lm, residual_thres, dx_thres, transform, dx_out,
                                               system_out, status);
            EGT_TRY(cudaGetLastError());
        }
    }
    return 0;
}

} // extern "C"
