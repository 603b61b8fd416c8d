// egt_tracking.cu -- image utilities of the dense tracker for sm_100a (include/eggtrack.h).
//
// Replaces the live kernels of /root/reference/src/utils/cuda/src/tracking.cu.  The reference launches 16x16 blocks
// that read every tap from global memory, re-uploads its stencil tables to __constant__ memory on every call and
// ends every wrapper with cudaDeviceSynchronize(); here the heavy stencil (13x13 bilateral, 169 expf per pixel) is
// tiled through shared memory, tables are compile-time constants, and everything is stream-ordered.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eggtrack.h"
#include "egt_qr.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ bilateral
constexpr int BL_TX = 32, BL_TY = 8, BL_MAXR = 8;

__global__ void __launch_bounds__(BL_TX * BL_TY)
k_bilateral(const float* __restrict__ in, float* __restrict__ out, int wd, int ht, int radius, float sc2inv,
            float ss2inv) {
    __shared__ float tile[(BL_TY + 2 * BL_MAXR)][(BL_TX + 2 * BL_MAXR) + 1];
    const int x0 = blockIdx.x * BL_TX - radius, y0 = blockIdx.y * BL_TY - radius;
    const int tw = BL_TX + 2 * radius, th = BL_TY + 2 * radius;
    for (int i = threadIdx.y * BL_TX + threadIdx.x; i < tw * th; i += BL_TX * BL_TY) {
        const int ly = i / tw, lx = i - ly * tw;
        const int gx = x0 + lx, gy = y0 + ly;
        tile[ly][lx] = (gx >= 0 && gx < wd && gy >= 0 && gy < ht) ? __ldg(in + (size_t)gy * wd + gx) : 0.f;
    }
    __syncthreads();
    const int x = blockIdx.x * BL_TX + threadIdx.x, y = blockIdx.y * BL_TY + threadIdx.y;
    if (x >= wd || y >= ht) return;
    const float center = tile[threadIdx.y + radius][threadIdx.x + radius];
    float sum1 = 0.f, sum2 = 0.f;
    for (int dy = -radius; dy <= radius; ++dy) {
        const int ny = y + dy;
        if (ny < 0 || ny >= ht) continue;
        for (int dx = -radius; dx <= radius; ++dx) {
            const int nx = x + dx;
            if (nx < 0 || nx >= wd) continue;
            const float v = tile[threadIdx.y + radius + dy][threadIdx.x + radius + dx];
            const float dc = center - v;
            const float space2 = (float)(dx * dx + dy * dy);
            const float w = expf(-space2 * ss2inv - dc * dc * sc2inv);
            sum1 += v * w;
            sum2 += w;
        }
    }
    out[(size_t)y * wd + x] = sum1 / sum2;
}

// generic fallback for windows larger than the tiled kernel supports
__global__ void k_bilateral_generic(const float* __restrict__ in, float* __restrict__ out, int wd, int ht, int radius,
                                    float sc2inv, float ss2inv) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wd || y >= ht) return;
    const float center = in[(size_t)y * wd + x];
    float sum1 = 0.f, sum2 = 0.f;
    for (int dy = -radius; dy <= radius; ++dy)
        for (int dx = -radius; dx <= radius; ++dx) {
            const int nx = x + dx, ny = y + dy;
            if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
            const float v = __ldg(in + (size_t)ny * wd + nx);
            const float dc = center - v;
            const float w = expf(-(float)(dx * dx + dy * dy) * ss2inv - dc * dc * sc2inv);
            sum1 += v * w;
            sum2 += w;
        }
    out[(size_t)y * wd + x] = sum1 / sum2;
}

// ------------------------------------------------------------------------------------------------ gaussian blur
__global__ void k_gaussian(const float* __restrict__ in, float* __restrict__ out, int wd, int ht, int ch, int radius,
                           float ss2inv) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wd || y >= ht) return;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2 = 0.f;
    for (int dy = -radius; dy <= radius; ++dy)
        for (int dx = -radius; dx <= radius; ++dx) {
            const int nx = x + dx, ny = y + dy;
            if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
            const float w = expf(-(float)(dx * dx + dy * dy) * ss2inv);
            for (int c = 0; c < ch; ++c) s1[c] += __ldg(in + ((size_t)ny * wd + nx) * ch + c) * w;
            s2 += w;
        }
    for (int c = 0; c < ch; ++c) out[((size_t)y * wd + x) * ch + c] = s1[c] / s2;
}

// ------------------------------------------------------------------------------------------------ pyramid step
__global__ void k_downsample(const float* __restrict__ in, float* __restrict__ out, int wd, int ht, int ch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int dw = wd / 2, dh = ht / 2;
    if (x >= dw || y >= dh) return;
    const float k1[5] = {1.f, 4.f, 6.f, 4.f, 1.f};
    float sum[4] = {0.f, 0.f, 0.f, 0.f}, count = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int nx = 2 * x + dx, ny = 2 * y + dy;
            if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
            const float w = k1[dy + 2] * k1[dx + 2];
            for (int c = 0; c < ch; ++c) sum[c] += __ldg(in + ((size_t)ny * wd + nx) * ch + c) * w;
            count += w;
        }
    for (int c = 0; c < ch; ++c) out[((size_t)y * dw + x) * ch + c] = sum[c] / count;
}

// ------------------------------------------------------------------------------------------------ gradients
__global__ void k_gradients(const float* __restrict__ in, float* __restrict__ gx, float* __restrict__ gy, int wd,
                            int ht) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wd || y >= ht) return;
    // tables of TRK:903-909; the reference walks them backwards (kernel_index 8 -> 0, TRK:869-889)
    const float kx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    const float ky[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
    float ax = 0.f, ay = 0.f;
    int k = 8;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int nx = x + dx, ny = y + dy;
            if (nx >= 0 && nx < wd && ny >= 0 && ny < ht) {
                const float v = __ldg(in + (size_t)ny * wd + nx);
                ax += v * kx[k];
                ay += v * ky[k];
            }
            --k;
        }
    gx[(size_t)y * wd + x] = ax;
    gy[(size_t)y * wd + x] = ay;
}

// ------------------------------------------------------------------------------------------------ vertex / normal
__device__ __forceinline__ float3 backproject(const float* __restrict__ depth, int wd, int x, int y, float fx, float fy,
                                              float cx, float cy) {
    const float Z = __ldg(depth + (size_t)y * wd + x);
    return make_float3((x - cx) * Z / fx, (y - cy) * Z / fy, Z);
}

// one pass: the reference runs two kernels with a device sync in between and re-reads the vertex map
__global__ void k_vertex_normal(const float* __restrict__ depth, float* __restrict__ vmap, float* __restrict__ nmap,
                                int wd, int ht, float fx, float fy, float cx, float cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= wd || y >= ht) return;
    const size_t idx = (size_t)y * wd + x;
    const float3 v00 = backproject(depth, wd, x, y, fx, fy, cx, cy);
    const float3 v10 = (x + 1 < wd) ? backproject(depth, wd, x + 1, y, fx, fy, cx, cy) : v00;
    const float3 v01 = (y + 1 < ht) ? backproject(depth, wd, x, y + 1, fx, fy, cx, cy) : v00;
    vmap[idx * 3] = v00.x; vmap[idx * 3 + 1] = v00.y; vmap[idx * 3 + 2] = v00.z;
    const float3 a = make_float3(v01.x - v00.x, v01.y - v00.y, v01.z - v00.z);
    const float3 b = make_float3(v10.x - v00.x, v10.y - v00.y, v10.z - v00.z);
    float3 n = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    const float inv = rsqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x *= inv; n.y *= inv; n.z *= inv;
    if (isnan(n.x) || isnan(n.y) || isnan(n.z)) n = make_float3(0.f, 0.f, 0.f);
    nmap[idx * 3] = n.x; nmap[idx * 3 + 1] = n.y; nmap[idx * 3 + 2] = n.z;
}

// ------------------------------------------------------------------------------------------------ small dense solve
// (A + lm I) x = b by the reference's own algorithm -- Eigen's column-pivoted Householder QR in fp32 on the column-major
// view of the buffer (egt_qr.cuh) -- but on the device, one thread: n <= 16 is 6 in practice (9 solves per frame); this
// removes the reference's GPU -> CPU (Eigen) -> GPU round trip and keeps its behaviour on non-symmetric and
// rank-deficient systems (basic solution, zeros for the dropped pivots).
__global__ void k_solve_block(const float* __restrict__ A, const float* __restrict__ b, float lm, float* __restrict__ x,
                              int n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    egt_colpiv_qr_solve(A, b, lm, x, n);
}

inline dim3 grid2(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }
#define EGT_TRY_LAUNCH()                                   \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)
} // namespace

extern "C" {

EGS_API int egt_bilateral_filter(const float* in, float* out, int32_t wd, int32_t ht, int32_t window, float sigma_c,
                                 float sigma_s, void* stream) {
    if (!in || !out || wd <= 0 || ht <= 0 || window <= 0) return EGS_E_BADARG;
    const float ss2inv = 1.0f / (2.0f * sigma_s * sigma_s), sc2inv = 1.0f / (2.0f * sigma_c * sigma_c);
    const int radius = window / 2;
    cudaStream_t s = (cudaStream_t)stream;
    if (radius <= BL_MAXR) {
        dim3 b(BL_TX, BL_TY);
        k_bilateral<<<grid2(wd, ht, b), b, 0, s>>>(in, out, wd, ht, radius, sc2inv, ss2inv);
    } else {
        dim3 b(16, 16);
        k_bilateral_generic<<<grid2(wd, ht, b), b, 0, s>>>(in, out, wd, ht, radius, sc2inv, ss2inv);
    }
    EGT_TRY_LAUNCH();
    return 0;
}

EGS_API int egt_gaussian_filter(const float* in, float* out, int32_t wd, int32_t ht, int32_t ch, int32_t window,
                                float sigma_s, void* stream) {
    if (!in || !out || wd <= 0 || ht <= 0 || window <= 0) return EGS_E_BADARG;
    if (ch < 1 || ch > 4) return EGS_E_UNSUPPORTED;
    dim3 b(32, 8);
    k_gaussian<<<grid2(wd, ht, b), b, 0, (cudaStream_t)stream>>>(in, out, wd, ht, ch, window / 2,
                                                                 1.0f / (2.0f * sigma_s * sigma_s));
    EGT_TRY_LAUNCH();
    return 0;
}

EGS_API int egt_gaussian_downsample(const float* in, float* out, int32_t wd, int32_t ht, int32_t ch, void* stream) {
    if (!in || !out || wd <= 0 || ht <= 0) return EGS_E_BADARG;
    if (ch < 1 || ch > 4) return EGS_E_UNSUPPORTED;
    if (wd / 2 == 0 || ht / 2 == 0) return 0;
    dim3 b(32, 8);
    k_downsample<<<grid2(wd / 2, ht / 2, b), b, 0, (cudaStream_t)stream>>>(in, out, wd, ht, ch);
    EGT_TRY_LAUNCH();
    return 0;
}

EGS_API int egt_compute_gradients(const float* in, float* gx, float* gy, int32_t wd, int32_t ht, void* stream) {
    if (!in || !gx || !gy || wd <= 0 || ht <= 0) return EGS_E_BADARG;
    dim3 b(32, 8);
    k_gradients<<<grid2(wd, ht, b), b, 0, (cudaStream_t)stream>>>(in, gx, gy, wd, ht);
    EGT_TRY_LAUNCH();
    return 0;
}

EGS_API int egt_vertex_normal_map(const float* depth, float fx, float fy, float cx, float cy, float* vmap, float* nmap,
                                  int32_t wd, int32_t ht, void* stream) {
    if (!depth || !vmap || !nmap || wd <= 0 || ht <= 0) return EGS_E_BADARG;
    dim3 b(32, 8);
    k_vertex_normal<<<grid2(wd, ht, b), b, 0, (cudaStream_t)stream>>>(depth, vmap, nmap, wd, ht, fx, fy, cx, cy);
    EGT_TRY_LAUNCH();
    return 0;
}

EGS_API int egt_solve_block(const float* A, const float* b, float lm, float* x, int32_t n, void* stream) {
    if (!A || !b || !x || n <= 0) return EGS_E_BADARG;
    if (n > 16) return EGS_E_UNSUPPORTED;
    k_solve_block<<<1, 32, 0, (cudaStream_t)stream>>>(A, b, lm, x, n);
    EGT_TRY_LAUNCH();
    return 0;
}

} // extern "C"
