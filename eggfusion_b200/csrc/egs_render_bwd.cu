// egs_render_bwd.cu -- reverse compositing walk: per-pixel gradients -> per-surfel screen-space gradients.
//
// Replaces renderCUDA<3> backward (DGS/cuda_rasterizer/backward.cu:419-676), which issues 13 float atomicAdds
// per contributing (pixel, surfel) pair.  This kernel is bound by instruction issue, not by HBM (ncu: DRAM < 2 %),
// so the design minimises warp-instructions per (tile, surfel):
//   * a warp owns an 8x4 pixel block.  The forward saved, per (instance, warp block), the 32-bit mask of pixels
//     that blended it; a warp walks only instances with a non-zero mask (ballot + find-first-set, back of the list
//     first) and each lane reads its bit instead of repeating the power / alpha / transmittance tests.  (Those
//     pairs are exactly the reference's `contributor < last_contributor && power <= 0 && alpha >= 1/255`: a
//     pixel is never `done` before its last contributor.)  With no decision left, exp() is one MUFU.EX2;
//   * the reference's 14 running accumulators (accum_rec / last_* for 3 colour, 3 normal, 1 depth channels) are
//     folded into ONE scalar per pixel.  With kappa_j = sum_ch feature_ch(j) * dL/dpixel_ch and
//     sigma_j = sum_{k behind j} w_k kappa_k, the reference's
//         dL/dalpha_j = T_j * sum_ch (c_ch - accum_ch) g_ch + [normalisation, opacity, background terms]
//     equals  T_j * kappa_j + (K0 - sigma_j) / (1 - alpha_j)  with a per-pixel constant K0 (same real-number value,
//     different rounding order; parity is checked at 1e-4 relative);
//   * the 13 partials of the 32 pixels of a warp are summed with a transposing butterfly (16 shuffles; lane 2v
//     ends up with the warp total of value v), warp totals are combined through shared memory, and each
//     (tile, surfel) leaves the CTA as four 16-byte vector reductions (red.global.add.v4.f32) into G[P][16];
//   * the walk starts at the CTA-wide maximum of n_contrib, skipping list tails nobody blended.
// Reference quirks kept: x10 on the per-surfel normal gradient only, un-weighted depth-differencing term on
// mean2D, conic.xy gradient not doubled (SURVEY 8 a-bis).
#include <stdlib.h>
#include "egs_common.cuh"

#define BWD_BATCH 64
#define BWD_WARPS (EGS_TILE_THREADS / 32)

__device__ __forceinline__ float conic_power_b(float cxx, float cxy, float cyy, float dx, float dy) {
    const float q = __fmaf_rn(__fmul_rn(cxx, dx), dx, __fmul_rn(__fmul_rn(cyy, dy), dy));
    const float dist = __fmaf_rn(__fmul_rn(__fmul_rn(2.f, cxy), dx), dy, q);
    return __fmul_rn(-0.5f, dist);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// After this, lane L holds in v[0] the sum over the warp of the callers' v[L >> 1].
__device__ __forceinline__ void warp_transpose_reduce16(float (&v)[16], int lane) {
    const unsigned full = 0xffffffffu;
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float send = hi ? v[i] : v[i + 8];
            const float keep = hi ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = hi ? v[i] : v[i + 4];
            const float keep = hi ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = hi ? v[i] : v[i + 2];
            const float keep = hi ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(full, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 2);
    }
    v[0] += __shfl_xor_sync(full, v[0], 1);
}

__global__ void __launch_bounds__(EGS_TILE_THREADS, 4)
k_render_backward(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec, ImgView im,
                  BinView bn, long long cap, const float* __restrict__ gC, const float* __restrict__ gN,
                  const float* __restrict__ gDp, const float* __restrict__ gOp, float* __restrict__ sg) {
    __shared__ float4 s_rec[BWD_BATCH * 4];
    __shared__ uint32_t s_id[BWD_BATCH];
    __shared__ __align__(16) uint32_t s_lm[BWD_BATCH * 8];
    __shared__ __align__(16) float s_part[BWD_WARPS][BWD_BATCH][16];
    __shared__ unsigned long long s_mask[BWD_WARPS];
    __shared__ int s_top;

    const int tile = blockIdx.x;
    const long long start = im.tile_offset[tile];
    long long end = im.tile_offset[tile + 1];
    if (end > cap) end = cap;
    const int n = (int)(end - start);
    if (n <= 0) return;

    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = tx * EGS_TILE + (warp & 1) * 8, by = ty * EGS_TILE + (warp >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)W * py + px;
    const uint32_t* __restrict__ plist = bn.point_list + start;
    const float pxf = (float)px, pyf = (float)py;

    if (threadIdx.x == 0) s_top = 0;
    __syncthreads();

    float T_final = 0.f, D_final = 0.f;
    int last_contributor = 0;
    float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gn0 = 0.f, gn1 = 0.f, gn2 = 0.f, gD = 0.f, gO = 0.f;
    if (inside) {
        T_final = im.final_T[pix];
        D_final = im.final_D[pix];
        last_contributor = (int)im.n_contrib[pix];
        gc0 = __ldg(gC + pix); gc1 = __ldg(gC + HW + pix); gc2 = __ldg(gC + 2 * HW + pix);
        gn0 = __ldg(gN + pix); gn1 = __ldg(gN + HW + pix); gn2 = __ldg(gN + 2 * HW + pix);
        gD = __ldg(gDp + pix);
        gO = __ldg(gOp + pix);
    }
    int warp_last = last_contributor;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, d));
    if (lane == 0 && warp_last > 0) atomicMax(&s_top, warp_last);
    __syncthreads();
    const int top0 = s_top;

    // per-pixel constants.  backward.cu:620-621: the depth image is D / (1 - T_final)
    const float one_m_Tf = 1.f - T_final;
    const float gDn = gD / one_m_Tf;
    const float bg_dot = __ldg(bg) * gc0 + __ldg(bg + 1) * gc1 + __ldg(bg + 2) * gc2;
    // K0 = [normalisation of depth] + [opacity image] - [background]   (backward.cu:621, :631, :638-641)
    const float K0 = gD * D_final / one_m_Tf / one_m_Tf * -T_final + T_final * (gO - bg_dot);
    const float kx = 2.f * 0.5f * (float)W, ky = 2.f * 0.5f * (float)H; // 2 * ddelx, 2 * ddely
    const float gn0x = gn0 * 10.f, gn1x = gn1 * 10.f, gn2x = gn2 * 10.f;   // backward.cu:604

    float T = T_final;
    float sigma = 0.f;
    const uint32_t rec_base = smem_addr(s_rec);

    for (int top = top0; top > 0; top -= BWD_BATCH) {
        const int m = min(BWD_BATCH, top);
        __syncthreads(); // previous batch fully combined before its staging buffers are reused
        if ((int)threadIdx.x < m) {
            const uint32_t id = __ldg(plist + (top - 1 - (int)threadIdx.x));
            s_id[threadIdx.x] = id;
            const float4* src = reinterpret_cast<const float4*>(rec + id);
            const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
            s_rec[threadIdx.x * 4] = a; s_rec[threadIdx.x * 4 + 1] = b;
            s_rec[threadIdx.x * 4 + 2] = c; s_rec[threadIdx.x * 4 + 3] = d;
            const uint4* lm = reinterpret_cast<const uint4*>(bn.lane_masks + 8 * (size_t)(start + top - 1 - (int)threadIdx.x));
            reinterpret_cast<uint4*>(s_lm)[2 * threadIdx.x] = __ldg(lm);
            reinterpret_cast<uint4*>(s_lm)[2 * threadIdx.x + 1] = __ldg(lm + 1);
        } else if (threadIdx.x < BWD_BATCH) {
            reinterpret_cast<uint4*>(s_lm)[2 * threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(s_lm)[2 * threadIdx.x + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();

        unsigned long long wmask = 0ull;
#pragma unroll
        for (int c = 0; c < BWD_BATCH / 32; c++) {
            // entry j of the batch is list position top-1-j (back to front)
            const int jl = c * 32 + lane;
            unsigned hits = __ballot_sync(0xffffffffu, s_lm[8 * jl + warp] != 0u);
            while (hits) {
                const int j = c * 32 + __ffs(hits) - 1;
                hits &= hits - 1;
                const bool act = (s_lm[8 * j + warp] >> lane) & 1u;
                const uint32_t rad = rec_base + 64u * (uint32_t)j;
                const float4 q0 = lds128(rad);
                const float4 q1 = lds128(rad + 16u);
                const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
                const float power = conic_power_b(q1.x, q1.y, q1.z, dx, dy);
                const float G = ex2_approx(power * 1.4426950408889634f);
                const float alpha = fminf(0.99f, q0.w * G);

                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = 0.f;
                if (act) {
                    const float4 q2 = lds128(rad + 32u), q3 = lds128(rad + 48u);
                    const float ra = rcp_approx(1.f - alpha);
                    T = T * ra;                       // transmittance in front of this splat
                    const float w = alpha * T;
                    const float d_cur = q1.w - (dx * q2.x + dy * q2.y);
                    // kappa = <features of this splat at this pixel, pixel gradients>
                    float kappa = q2.z * gc0;
                    kappa = fmaf(q2.w, gc1, kappa); kappa = fmaf(q3.x, gc2, kappa);
                    kappa = fmaf(q3.y, gn0, kappa); kappa = fmaf(q3.z, gn1, kappa); kappa = fmaf(q3.w, gn2, kappa);
                    kappa = fmaf(d_cur, gDn, kappa);
                    const float dL_dalpha = fmaf(T, kappa, ra * (K0 - sigma));
                    sigma = fmaf(w, kappa, sigma);
                    const float dL_ddist = dL_dalpha * (q0.w * -0.5f * G);
                    v[0] = fmaf(dL_ddist * kx, q1.x * dx + q1.y * dy, -gD * q2.x);   // backward.cu:648-660
                    v[1] = fmaf(dL_ddist * ky, q1.z * dy + q1.y * dx, -gD * q2.y);
                    v[2] = dL_ddist * dx * dx;
                    v[3] = dL_ddist * dx * dy;
                    v[4] = dL_ddist * dy * dy;
                    v[5] = G * dL_dalpha;
                    v[6] = w * gc0; v[7] = w * gc1; v[8] = w * gc2;
                    v[9] = w * gn0x; v[10] = w * gn1x; v[11] = w * gn2x;
                    v[12] = w * gDn;
                }
                warp_transpose_reduce16(v, lane);
                if (!(lane & 1)) s_part[warp][j][lane >> 1] = v[0];
                wmask |= 1ull << j;
            }
        }
        if (lane == 0) s_mask[warp] = wmask;
        __syncthreads();

        // combine the warps' totals: thread -> (record, quad); one 16-byte vector reduction per quad
        {
            const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
            if (r < m) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                bool any = false;
#pragma unroll
                for (int wv = 0; wv < BWD_WARPS; wv++) {
                    if (s_mask[wv] >> r & 1ull) {
                        const float4 p = *reinterpret_cast<const float4*>(&s_part[wv][r][4 * q]);
                        s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
                        any = true;
                    }
                }
                if (any) red_add_v4(sg + (size_t)EGS_SCREEN_GRAD_STRIDE * s_id[r] + 4 * q, s.x, s.y, s.z, s.w);
            }
        }
    }
}

cudaError_t launch_render_backward_mma(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                       const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                       cudaStream_t s);
cudaError_t launch_render_backward_warp(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                        const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                        cudaStream_t s);
cudaError_t launch_render_backward_gather(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                          const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                          cudaStream_t s);

// Four implementations of the same walk, selected with EGS_BWD_KERNEL (default: warp).  Measured on B200 at C3:
//   warp       one 8x4 block per 32-thread CTA, per-warp compacted staging, no CTA barriers
//              (egs_render_bwd_warp.cu)                                                                  1.00 ms
//   gather     one tile per CTA, transposed FP32 reduction + cp.async double-buffered staging
//              (egs_render_bwd_gather.cu)                                                                1.04 ms
//   butterfly  16-shuffle transposing butterfly + shared-memory combine (this file)                      1.38 ms
//   mma        cross-pixel sums on the tensor cores, mma.sync 3xTF32 (egs_render_bwd_mma.cu)             1.62 ms
// The legacy mma.sync path costs more issue slots than it saves here.  All four are covered by the parity tests.
//   lane       like warp, but the cross-pixel sums are taken by "a lane owns a splat" over 32-hit chunks with packed
//              FP32 and broadcast weight reads (egs_render_bwd_lane.cu): default since round 2
// Variants 3 (warp, lane) consume the forward's per-block hit lists, 0-2 its lane_masks.
cudaError_t launch_render_backward_lane(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                        const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                        cudaStream_t s);
static int g_bwd_lane = 1;
int egs_bwd_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("EGS_BWD_KERNEL");
        v = (e && e[0] == 'm') ? 1 : (e && e[0] == 'b') ? 0 : (e && e[0] == 'g') ? 2 : 3;   // default: lane
        g_bwd_lane = (e && e[0] == 'w') ? 0 : 1;
    }
    return v;
}

cudaError_t launch_render_backward(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                   const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                   cudaStream_t s) {
    if (egs_bwd_variant() == 1) return launch_render_backward_mma(f, g, im, bn, cap, gC, gN, gD, gO, sg, s);
    if (egs_bwd_variant() == 3)
        return g_bwd_lane ? launch_render_backward_lane(f, g, im, bn, cap, gC, gN, gD, gO, sg, s)
                          : launch_render_backward_warp(f, g, im, bn, cap, gC, gN, gD, gO, sg, s);
    if (egs_bwd_variant() == 2) return launch_render_backward_gather(f, g, im, bn, cap, gC, gN, gD, gO, sg, s);
    const int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
    k_render_backward<<<gx * gy, EGS_TILE_THREADS, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap, gC, gN, gD,
                                                           gO, sg);
    return cudaGetLastError();
}
