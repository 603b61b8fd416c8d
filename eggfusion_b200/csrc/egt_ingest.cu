// egt_ingest.cu -- frame ingest of the dense tracker as one stream-ordered chain of fused, shared-memory-tiled kernels
// (C ABI: egt_ingest_frame in include/eggtrack.h).
//
// The reference builds the pyramids of a frame (and, per frame, of the rendered model map) in Python
// (/root/reference/src/utils/frame.py:112-146 Frame.__init__, :32-99 PyraImageCUDA): per frame 3 bilateral filters,
// ~14 stencil launches of src/utils/cuda/src/tracking.cu:533-926 (each ending in cudaDeviceSynchronize and re-uploading
// its tables to __constant__ memory), and ~25 elementwise torch launches between them.  Here it is nlevel launches:
//   k_ingest_level0   raw depth -> 13x13 bilateral (tile + halo staged in shared memory) -> depth / disparity / mask,
//                     vertex + normal map (the filtered tile is reused for the x+1 / y+1 neighbours), grey image and its
//                     3x3 derivative + magnitude (grey tile with a 1-pixel halo in shared memory)
//   k_ingest_down     one pyramid step: 5x5 binomial stride-2 downsample of grey / depth / mask / vertex / normal, the
//                     13x13 bilateral of the downsampled depth (downsampled tile + halo kept in shared memory, never
//                     written out unfiltered), normalisation of the normal, derivative of the downsampled grey
// Every stage keeps the arithmetic of the per-function kernels in egt_tracking.cu (which are pinned on the reference's
// own kernels, tests/golden/tracking_161x119.npz): same tap order, same border rules (out-of-image taps skipped and the
// weights renormalised, one IEEE expf per bilateral tap), torch's elementwise expressions written with explicit
// roundings -- the chain reproduces the per-function sequence bit for bit except for FMA contraction inside the
// stencils.  (A faster bilateral -- tabulated spatial weight times __expf of the range term -- was measured and
// rejected: the filtered depth moved by 3e-7 relative, which the normals' cross products of neighbouring depth
// differences amplify to 6e-5; 169 expf per pixel keep the chain SFU / issue bound rather than HBM bound.)
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eggtrack.h"

namespace {
constexpr int TX = 32, TY = 8;          // output tile (one thread per pixel)
constexpr int BR = 6;                   // radius of the 13x13 bilateral of frame.py:84,132

__device__ __forceinline__ float gray_of(const float* __restrict__ c) {
    // (c0 * 0.114 + c1 * 0.587) + c2 * 0.299 with torch's separate roundings (frame.py:40; RGB_COEFF is read backwards)
    return __fadd_rn(__fadd_rn(__fmul_rn(c[0], 0.114f), __fmul_rn(c[1], 0.587f)), __fmul_rn(c[2], 0.299f));
}

// 13x13 bilateral at tile position (lx, ly) of a staged tile whose element (0,0) is image pixel (x0, y0)
template <int PITCH>
__device__ __forceinline__ float bilateral_at(const float (*tile)[PITCH], int lx, int ly, int x, int y, int wd, int ht,
                                              float sc2inv, float ss2inv) {
    const float center = tile[ly][lx];
    float sum1 = 0.f, sum2 = 0.f;
    for (int dy = -BR; dy <= BR; ++dy) {
        const int ny = y + dy;
        if (ny < 0 || ny >= ht) continue;
#pragma unroll
        for (int dx = -BR; dx <= BR; ++dx) {
            const int nx = x + dx;
            if (nx < 0 || nx >= wd) continue;
            const float v = tile[ly + dy][lx + dx];
            const float dc = center - v;
            const float space2 = (float)(dx * dx + dy * dy);
            const float w = expf(-space2 * ss2inv - dc * dc * sc2inv);
            sum1 += v * w;
            sum2 += w;
        }
    }
    return sum1 / sum2;
}

// 3x3 derivative pair of egt_tracking.cu:k_gradients on a grey tile with a 1-pixel halo (element (0,0) = pixel (x-1, y-1)
// of the tile origin)
template <int PITCH>
__device__ __forceinline__ void gradient_at(const float (*g)[PITCH], int lx, int ly, int x, int y, int wd, int ht, float& ax,
                                            float& ay) {
    const float kx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    const float ky[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
    ax = 0.f;
    ay = 0.f;
    int k = 8;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int nx = x + dx, ny = y + dy;
            if (nx >= 0 && nx < wd && ny >= 0 && ny < ht) {
                const float v = g[ly + 1 + dy][lx + 1 + dx];
                ax += v * kx[k];
                ay += v * ky[k];
            }
            --k;
        }
}

__device__ __forceinline__ void store_grad(float* __restrict__ grad, size_t idx, float ax, float ay) {
    grad[3 * idx] = ax;
    grad[3 * idx + 1] = ay;
    grad[3 * idx + 2] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), 1e-6f));   // frame.py:73
}

// ---- level 0 -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TX * TY)
k_ingest_level0(const float* __restrict__ color, const float* __restrict__ depth_raw, const float* __restrict__ maskf,
                int wd, int ht, float fx, float fy, float cx, float cy, float sc2inv, float ss2inv, egt_pyramid_level o) {
    __shared__ float raw[TY + 1 + 2 * BR][TX + 1 + 2 * BR + 1];   // raw depth: tile + 1 (x+1 / y+1 neighbours) + bilateral halo
    __shared__ float filt[TY + 1][TX + 1 + 1];                    // filtered depth of the tile + its +1 column / row
    __shared__ float gry[TY + 2][TX + 2 + 1];                     // grey: tile + 1-pixel halo
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int i = tid; i < (TY + 1 + 2 * BR) * (TX + 1 + 2 * BR); i += TX * TY) {
        const int ly = i / (TX + 1 + 2 * BR), lx = i - ly * (TX + 1 + 2 * BR);
        const int gx = tx0 - BR + lx, gy = ty0 - BR + ly;
        raw[ly][lx] = (gx >= 0 && gx < wd && gy >= 0 && gy < ht) ? __ldg(depth_raw + (size_t)gy * wd + gx) : 0.f;
    }
    for (int i = tid; i < (TY + 2) * (TX + 2); i += TX * TY) {
        const int ly = i / (TX + 2), lx = i - ly * (TX + 2);
        const int gx = tx0 - 1 + lx, gy = ty0 - 1 + ly;
        gry[ly][lx] = (gx >= 0 && gx < wd && gy >= 0 && gy < ht) ? gray_of(color + 3 * ((size_t)gy * wd + gx)) : 0.f;
    }
    __syncthreads();
    // bilateral for the tile and for its +1 column / row (frame.py:132: cuda_bilateral_filter(depth, 13, 0.03, 4.5))
    for (int i = tid; i < (TY + 1) * (TX + 1); i += TX * TY) {
        const int ly = i / (TX + 1), lx = i - ly * (TX + 1);
        const int x = tx0 + lx, y = ty0 + ly;
        filt[ly][lx] = (x < wd && y < ht) ? bilateral_at(raw, lx + BR, ly + BR, x, y, wd, ht, sc2inv, ss2inv) : 0.f;
    }
    __syncthreads();
    const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
    if (x >= wd || y >= ht) return;
    const size_t idx = (size_t)y * wd + x;
    const int lx = threadIdx.x, ly = threadIdx.y;
    const float Z = filt[ly][lx];
    o.depth[idx] = Z;
    o.disp[idx] = __fdiv_rn(1.0f, __fadd_rn(Z, 1e-6f));                      // frame.py:68
    const float mf = __ldg(maskf + idx);
    o.maskf[idx] = mf;
    o.mask[idx] = (uint8_t)(mf > 0.9f && Z > 0.1f);                          // frame.py:69
    // vertex + normal (egt_tracking.cu:k_vertex_normal; the reference: tracking.cu:602-702)
    const float3 v00 = make_float3((x - cx) * Z / fx, (y - cy) * Z / fy, Z);
    float3 v10 = v00, v01 = v00;
    if (x + 1 < wd) { const float Zr = filt[ly][lx + 1]; v10 = make_float3((x + 1 - cx) * Zr / fx, (y - cy) * Zr / fy, Zr); }
    if (y + 1 < ht) { const float Zd = filt[ly + 1][lx]; v01 = make_float3((x - cx) * Zd / fx, (y + 1 - cy) * Zd / fy, Zd); }
    o.vertex[idx * 3] = v00.x; o.vertex[idx * 3 + 1] = v00.y; o.vertex[idx * 3 + 2] = v00.z;
    const float3 a = make_float3(v01.x - v00.x, v01.y - v00.y, v01.z - v00.z);
    const float3 b = make_float3(v10.x - v00.x, v10.y - v00.y, v10.z - v00.z);
    float3 n = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    const float inv = rsqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x *= inv; n.y *= inv; n.z *= inv;
    if (isnan(n.x) || isnan(n.y) || isnan(n.z)) n = make_float3(0.f, 0.f, 0.f);
    o.normal[idx * 3] = n.x; o.normal[idx * 3 + 1] = n.y; o.normal[idx * 3 + 2] = n.z;
    o.gray[idx] = gry[ly + 1][lx + 1];
    float ax, ay;
    gradient_at(gry, lx, ly, x, y, wd, ht, ax, ay);
    store_grad(o.grad, idx, ax, ay);
}

// ---- one pyramid step --------------------------------------------------------------------------------------------------
// 5x5 binomial at stride 2 of channel c of a [h][w][ch] image, normalised by the in-image weight (k_downsample)
__device__ __forceinline__ float down_tap(const float* __restrict__ in, int wd, int ht, int ch, int c, int x, int y) {
    const float k1[5] = {1.f, 4.f, 6.f, 4.f, 1.f};
    float sum = 0.f, count = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int nx = 2 * x + dx, ny = 2 * y + dy;
            if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
            const float w = k1[dy + 2] * k1[dx + 2];
            sum += __ldg(in + ((size_t)ny * wd + nx) * ch + c) * w;
            count += w;
        }
    return sum / count;
}

__global__ void __launch_bounds__(TX * TY)
k_ingest_down(egt_pyramid_level p, egt_pyramid_level o, float sc2inv, float ss2inv) {
    __shared__ float dd[TY + 2 * BR][TX + 2 * BR + 1];   // downsampled depth: tile + bilateral halo
    __shared__ float gry[TY + 2][TX + 2 + 1];            // downsampled grey: tile + 1-pixel halo
    const int wd = p.width, ht = p.height, dw = o.width, dh = o.height;
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int i = tid; i < (TY + 2 * BR) * (TX + 2 * BR); i += TX * TY) {
        const int ly = i / (TX + 2 * BR), lx = i - ly * (TX + 2 * BR);
        const int gx = tx0 - BR + lx, gy = ty0 - BR + ly;
        dd[ly][lx] = (gx >= 0 && gx < dw && gy >= 0 && gy < dh) ? down_tap(p.depth, wd, ht, 1, 0, gx, gy) : 0.f;   // frame.py:83
    }
    for (int i = tid; i < (TY + 2) * (TX + 2); i += TX * TY) {
        const int ly = i / (TX + 2), lx = i - ly * (TX + 2);
        const int gx = tx0 - 1 + lx, gy = ty0 - 1 + ly;
        gry[ly][lx] = (gx >= 0 && gx < dw && gy >= 0 && gy < dh) ? down_tap(p.gray, wd, ht, 1, 0, gx, gy) : 0.f;    // frame.py:77
    }
    __syncthreads();
    const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const size_t idx = (size_t)y * dw + x;
    const int lx = threadIdx.x, ly = threadIdx.y;
    const float Z = bilateral_at(dd, lx + BR, ly + BR, x, y, dw, dh, sc2inv, ss2inv);                                   // frame.py:84
    o.depth[idx] = Z;
    o.disp[idx] = __fdiv_rn(1.0f, __fadd_rn(Z, 1e-6f));
    const float mf = down_tap(p.maskf, wd, ht, 1, 0, x, y);                                                          // frame.py:87
    o.maskf[idx] = mf;
    o.mask[idx] = (uint8_t)(mf > 0.9f && Z > 0.1f);
    float n[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        o.vertex[idx * 3 + c] = down_tap(p.vertex, wd, ht, 3, c, x, y);                                              // frame.py:90
        n[c] = down_tap(p.normal, wd, ht, 3, c, x, y);                                                               // frame.py:93
    }
    // F.normalize(normal, dim=-1): x / max(||x||_2, 1e-12)                                                          // frame.py:94
    const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2]))), 1e-12f);
#pragma unroll
    for (int c = 0; c < 3; c++) o.normal[idx * 3 + c] = __fdiv_rn(n[c], nrm);
    o.gray[idx] = gry[ly + 1][lx + 1];
    float ax, ay;
    gradient_at(gry, lx, ly, x, y, dw, dh, ax, ay);
    store_grad(o.grad, idx, ax, ay);
}
} // namespace

extern "C" EGS_API int egt_ingest_frame(const float* color, const float* depth_raw, const float* mask, int32_t width,
                                        int32_t height, float fx, float fy, float cx, float cy, float sigma_color,
                                        float sigma_space, int32_t nlevel, const egt_pyramid_level* levels, void* stream) {
    if (!color || !depth_raw || !mask || !levels || width <= 0 || height <= 0 || nlevel < 1 || nlevel > 8) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const float ss2inv = 1.0f / (2.0f * sigma_space * sigma_space), sc2inv = 1.0f / (2.0f * sigma_color * sigma_color);
    int w = width, h = height;
    for (int l = 0; l < nlevel; l++) {
        const egt_pyramid_level& o = levels[l];
        if (o.width != w || o.height != h) return EGS_E_BADARG;
        if (!o.depth || !o.disp || !o.mask || !o.maskf || !o.vertex || !o.normal || !o.gray || !o.grad) return EGS_E_BADARG;
        const dim3 b(TX, TY), g((w + TX - 1) / TX, (h + TY - 1) / TY);
        if (l == 0)
            k_ingest_level0<<<g, b, 0, s>>>(color, depth_raw, mask, w, h, fx, fy, cx, cy, sc2inv, ss2inv, o);
        else
            k_ingest_down<<<g, b, 0, s>>>(levels[l - 1], o, sc2inv, ss2inv);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        w /= 2;
        h /= 2;
        if ((w == 0 || h == 0) && l + 1 < nlevel) return EGS_E_BADARG;
    }
    return 0;
}
