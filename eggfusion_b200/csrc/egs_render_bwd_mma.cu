// egs_render_bwd_mma.cu -- reverse compositing walk, variant whose cross-pixel sums run on the tensor cores.
//
// Same per-pixel arithmetic as egs_render_bwd.cu (see there for the algebra and the reference quirks).  What
// differs is how the 32 pixels of a warp are summed per splat.  Every per-surfel partial is a FIXED per-pixel
// weight times a per-(pixel, splat) scalar:
//     colour/normal/depth grads   = sum_px K_i(px) * w(px)            K = (gC, 10 gN, gD/(1-Tf)),   w = alpha*T
//     conic / mean2D grads        = polynomials in the splat centre of the moments  sum_px X^a Y^b * dd(px)
//                                    (X, Y = pixel offsets from the tile centre, dd = dL/d(dist))
//     opacity grad                = sum_px 1 * u(px)                   u = G * dL/dalpha
//     depth-differencing term     = -j_{a,b} * sum_px gD(px) * act(px)
// so two splats at a time are reduced as D[16x8] = A[16x32] * B[32x8]: A = the warp's constant weight rows
// (7 K rows, gD, 6 moment rows), B = (w, dd, u, act) x 2 splats, k = the 32 pixels.  mma.sync.m16n8k8 TF32 with
// the 3xTF32 split (A_hi B_hi + A_lo B_hi + A_hi B_lo; moment rows are small half-integers, exact in TF32) keeps
// fp32-level accuracy.  ~36 issue slots per (warp, splat) instead of ~76 for the shuffle butterfly.
// The moments are turned into conic / mean gradients once per (tile, splat), after the sum over the 8 warps.
#include "egs_common.cuh"

#define BWD_BATCH 64
#define BWD_WARPS (EGS_TILE_THREADS / 32)
#define SLAB_STRIDE 36 // floats per B column in shared memory: conflict-free fragment loads (4g + t)

namespace {

__device__ __forceinline__ float conic_power_m(float cxx, float cxy, float cyy, float dx, float dy) {
    const float q = __fmaf_rn(__fmul_rn(cxx, dx), dx, __fmul_rn(__fmul_rn(cyy, dy), dy));
    const float dist = __fmaf_rn(__fmul_rn(__fmul_rn(2.f, cxy), dx), dy, q);
    return __fmul_rn(-0.5f, dist);
}
__device__ __forceinline__ float ex2_approx_m(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx_m(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_add_v4_m(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// x = hi + lo with hi, lo representable in TF32 (lo carries the next 11 mantissa bits)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
// D += A(16x8, row) * B(8x8, col), TF32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

} // namespace

// s_part slots per (warp, record): 0..5 moments M00 M10 M01 M20 M11 M02 | 6..12 colour(3) normal(3) depth | 13 sum u |
// 14 sum gD*act | 15 unused
__global__ void __launch_bounds__(EGS_TILE_THREADS, 3)
k_render_backward_mma(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec,
                      ImgView im, BinView bn, long long cap, const float* __restrict__ gC,
                      const float* __restrict__ gN, const float* __restrict__ gDp, const float* __restrict__ gOp,
                      float* __restrict__ sg) {
    __shared__ float4 s_rec[BWD_BATCH * 4];
    __shared__ uint32_t s_id[BWD_BATCH];
    __shared__ __align__(16) uint32_t s_lm[BWD_BATCH * 8];
    __shared__ __align__(16) float s_part[BWD_WARPS][BWD_BATCH][16];
    __shared__ __align__(16) float s_slab[BWD_WARPS][8 * SLAB_STRIDE];
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
    const float tile_cx = (float)(tx * EGS_TILE) + 7.5f, tile_cy = (float)(ty * EGS_TILE) + 7.5f;

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

    const float one_m_Tf = 1.f - T_final;
    const float gDn = gD / one_m_Tf;
    const float bg_dot = __ldg(bg) * gc0 + __ldg(bg + 1) * gc1 + __ldg(bg + 2) * gc2;
    const float K0 = gD * D_final / one_m_Tf / one_m_Tf * -T_final + T_final * (gO - bg_dot);
    const float kx = 2.f * 0.5f * (float)W, ky = 2.f * 0.5f * (float)H;

    // ---- A fragments (constant for the whole walk).  Row r of A, pixel k:
    //   r = 0..6 : K_r(k) = gc0 gc1 gc2 10gn0 10gn1 10gn2 gDn      r = 7 : gD(k)       r = 8..13 : 1 X Y XX XY YY
    // m16n8k8 layout: lane (g = lane>>2, t = lane&3) holds a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4],
    // a3 = A[g+8][t+4] of every 8-pixel k-step s (pixel k = 8 s + column).
    const int g = lane >> 2, t = lane & 3;
    uint32_t a_hi0[4], a_hi2[4], a_lo0[4], a_lo2[4], a_m1[4], a_m3[4];
    {
        float* tab = &s_part[warp][0][0]; // scratch: [8 rows][32 pixels]
        tab[0 * 32 + lane] = gc0; tab[1 * 32 + lane] = gc1; tab[2 * 32 + lane] = gc2;
        tab[3 * 32 + lane] = gn0 * 10.f; tab[4 * 32 + lane] = gn1 * 10.f; tab[5 * 32 + lane] = gn2 * 10.f;
        tab[6 * 32 + lane] = gDn; tab[7 * 32 + lane] = gD;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; s++) {
            split_tf32(tab[g * 32 + 8 * s + t], a_hi0[s], a_lo0[s]);
            split_tf32(tab[g * 32 + 8 * s + t + 4], a_hi2[s], a_lo2[s]);
            // moment rows: pixel k = 8 s + c sits at block-local (c, s); offsets from the tile centre
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int c = t + 4 * h;
                const float X = (float)((warp & 1) * 8 + c) - 7.5f, Y = (float)((warp >> 1) * 4 + s) - 7.5f;
                const float mv = g == 0 ? 1.f : g == 1 ? X : g == 2 ? Y : g == 3 ? X * X : g == 4 ? X * Y : g == 5 ? Y * Y : 0.f;
                (h == 0 ? a_m1[s] : a_m3[s]) = to_tf32(mv); // exact: |value| <= 56.25 in steps of 0.25
            }
        }
    }
    __syncthreads();
    const int top0 = s_top;

    float T = T_final;
    float sigma = 0.f;
    const uint32_t rec_base = smem_addr(s_rec);
    float* slab = s_slab[warp];

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
        int npend = 0, jpend = 0;

        // reduce the (up to two) pending splats of the slab: D = A * B, then scatter the useful entries of D
        auto flush = [&](int j0, int j1) {
            __syncwarp();
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(slab[g * SLAB_STRIDE + 8 * s + t], bh0, bl0);
                split_tf32(slab[g * SLAB_STRIDE + 8 * s + t + 4], bh1, bl1);
                mma_tf32(d0, a_hi0[s], a_m1[s], a_hi2[s], a_m3[s], bh0, bh1);
                mma_tf32(d1, a_lo0[s], 0u, a_lo2[s], 0u, bh0, bh1);
                mma_tf32(d2, a_hi0[s], a_m1[s], a_hi2[s], a_m3[s], bl0, bl1);
            }
            // D[row][col]: c0 = (g, 2t)  c1 = (g, 2t+1)  c2 = (g+8, 2t)  c3 = (g+8, 2t+1);  col = 4*splat + type
            const float c0 = d0[0] + d1[0] + d2[0], c1 = d0[1] + d1[1] + d2[1];
            const float c2 = d0[2] + d1[2] + d2[2], c3 = d0[3] + d1[3] + d2[3];
            const int j = (t & 2) ? j1 : j0;
            if (j >= 0) {
                float* out = &s_part[warp][j][0];
                if (!(t & 1)) {          // columns (w, dd)
                    if (g <= 6) out[6 + g] = c0;   // K rows x w
                    if (g <= 5) out[g] = c3;       // moment rows x dd
                } else {                 // columns (u, act)
                    if (g == 7) out[14] = c1;      // gD row x act
                    if (g == 0) out[13] = c2;      // ones row x u
                }
            }
            __syncwarp();
        };

#pragma unroll 1
        for (int c = 0; c < BWD_BATCH / 32; c++) {
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
                const float power = conic_power_m(q1.x, q1.y, q1.z, dx, dy);
                const float G = ex2_approx_m(power * 1.4426950408889634f);
                const float alpha = fminf(0.99f, q0.w * G);
                float w = 0.f, dd = 0.f, u = 0.f;
                if (act) {
                    const float4 q2 = lds128(rad + 32u), q3 = lds128(rad + 48u);
                    const float ra = rcp_approx_m(1.f - alpha);
                    T = T * ra;
                    w = alpha * T;
                    const float d_cur = q1.w - (dx * q2.x + dy * q2.y);
                    float kappa = q2.z * gc0;
                    kappa = fmaf(q2.w, gc1, kappa); kappa = fmaf(q3.x, gc2, kappa);
                    kappa = fmaf(q3.y, gn0, kappa); kappa = fmaf(q3.z, gn1, kappa); kappa = fmaf(q3.w, gn2, kappa);
                    kappa = fmaf(d_cur, gDn, kappa);
                    const float dL_dalpha = fmaf(T, kappa, ra * (K0 - sigma));
                    sigma = fmaf(w, kappa, sigma);
                    dd = dL_dalpha * (q0.w * -0.5f * G);
                    u = G * dL_dalpha;
                }
                float* col = slab + (4 * npend) * SLAB_STRIDE + lane;
                col[0] = w; col[SLAB_STRIDE] = dd; col[2 * SLAB_STRIDE] = u; col[3 * SLAB_STRIDE] = act ? 1.f : 0.f;
                wmask |= 1ull << j;
                if (npend == 0) { jpend = j; npend = 1; }
                else { flush(jpend, j); npend = 0; }
            }
        }
        if (npend == 1) {
            float* col = slab + 4 * SLAB_STRIDE + lane;
            col[0] = 0.f; col[SLAB_STRIDE] = 0.f; col[2 * SLAB_STRIDE] = 0.f; col[3 * SLAB_STRIDE] = 0.f;
            flush(jpend, -1);
        }
        if (lane == 0) s_mask[warp] = wmask;
        __syncthreads();

        // combine the warps' raw sums per record, turn moments into conic / mean gradients, one vector reduction per quad
        if (threadIdx.x < (unsigned)m) {
            const int r = threadIdx.x;
            float raw[16];
#pragma unroll
            for (int i = 0; i < 16; i++) raw[i] = 0.f;
            bool any = false;
#pragma unroll
            for (int wv = 0; wv < BWD_WARPS; wv++) {
                if (s_mask[wv] >> r & 1ull) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float4 p = *reinterpret_cast<const float4*>(&s_part[wv][r][4 * q]);
                        raw[4 * q] += p.x; raw[4 * q + 1] += p.y; raw[4 * q + 2] += p.z; raw[4 * q + 3] += p.w;
                    }
                    any = true;
                }
            }
            if (any) {
                const float4 q0 = s_rec[4 * r], q1 = s_rec[4 * r + 1], q2 = s_rec[4 * r + 2];
                const float xc = q0.x - tile_cx, yc = q0.y - tile_cy;   // splat centre relative to the tile centre
                const float M00 = raw[0], M10 = raw[1], M01 = raw[2], M20 = raw[3], M11 = raw[4], M02 = raw[5];
                const float Sdx = xc * M00 - M10, Sdy = yc * M00 - M01; // sum dd*dx, sum dd*dy  (dx = xc - X)
                const float Gs = raw[14];
                const float v0 = kx * (q1.x * Sdx + q1.y * Sdy) - q2.x * Gs;
                const float v1 = ky * (q1.z * Sdy + q1.y * Sdx) - q2.y * Gs;
                const float v2 = xc * (xc * M00 - 2.f * M10) + M20;
                const float v3 = xc * (yc * M00 - M01) - yc * M10 + M11;
                const float v4 = yc * (yc * M00 - 2.f * M01) + M02;
                float* dst = sg + (size_t)EGS_SCREEN_GRAD_STRIDE * s_id[r];
                red_add_v4_m(dst, v0, v1, v2, v3);
                red_add_v4_m(dst + 4, v4, raw[13], raw[6], raw[7]);
                red_add_v4_m(dst + 8, raw[8], raw[9], raw[10], raw[11]);
                red_add_v4_m(dst + 12, raw[12], 0.f, 0.f, 0.f);
            }
        }
    }
}

cudaError_t launch_render_backward_mma(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                       const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                       cudaStream_t s) {
    const int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
    k_render_backward_mma<<<gx * gy, EGS_TILE_THREADS, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap, gC, gN,
                                                               gD, gO, sg);
    return cudaGetLastError();
}
