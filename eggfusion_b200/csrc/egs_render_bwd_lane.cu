// egs_render_bwd_lane.cu -- reverse compositing walk, one warp (one 8x4 pixel block) per CTA, cross-pixel sums by
// "a lane owns a splat".
//
// Same algebra, quirks and inputs as egs_render_bwd_warp.cu (reference: renderCUDA backward, DGS/cuda_rasterizer/
// backward.cu:419-676; the sigma form is derived in egs_render_bwd.cu): the warp walks its block's hit list
// {surfel id, pixel mask} back to front in chunks of 32 hits whose 64-byte records are staged by the TMA engine
// (one cp.async.bulk per record, completing on an mbarrier; SASS UBLKCP / SYNCS), double-buffered.
// What changes is the cross-pixel reduction.  ncu on the warp variant (round 2, profiles/README.md): the LSU data pipe
// is 95 % busy -- 30 shared-memory wavefronts per (block, splat) hit, 12 of them the phase-2 gather (lane = splat x
// pixel-quarter re-reading parked pairs and a per-pixel weight table whose 16-byte entries differ across the warp: 4
// wavefronts per LDS.128) -- so the kernel is bound by shared-memory wavefronts, not by issue slots or HBM.  Here:
//   * the records arrive through the TMA engine instead of cp.async: ncu's source page showed every 16-byte LDGSTS of a
//     gathered record costing one shared-memory wavefront PER LANE (32 per instruction, 5.4 per hit, 21 % of the
//     kernel's wavefronts); a bulk copy writes shared memory through the async proxy and leaves the LSU pipe alone.
//     Measured at C3 (bwd_render stage): cp.async 0.712 ms, TMA 0.682 ms, LDG.128 into registers one chunk ahead +
//     conflict-free STS.128 into a quad-major layout 0.716 ms (the scattered global loads cost the LSU pipe what the
//     LDGSTS did); the per-lane bulk copies are serialised through uniform registers (~8 issue slots each);
//   * phase 1 (lane = pixel) parks only {w, u} per (pixel, splat) -- dd = u * (-opacity/2) is a per-splat factor applied
//     to the finished sums, and "this pixel blended the splat" is recovered as w > 0 -- as two conflict-free 4-byte
//     stores (row = 32 w | 32 u): 2 wavefronts per hit instead of 4;
//   * phase 2 runs once per 32-hit chunk with lane = splat: every lane walks the 16 pixel pairs of ITS splat's parked
//     row (two conflict-free LDS.64 per pair) against the block's weight table, which all lanes now read at the same
//     address (a broadcast: 2 wavefronts per LDS.128, measured).  The 14 sums of two horizontally adjacent pixels are accumulated
//     with packed FP32 (FFMA2 / FMUL2 / FADD2 with scalar-broadcast operands), folded once at the end, and the lane
//     issues its splat's four red.global.add.v4.f32 itself: no shuffles, no second staging of the records, 4 instead
//     of 12 phase-2 wavefronts and ~13 instead of ~30 phase-2 instructions per hit.
// A chunk is exactly one phase-2 pass, so the parked rows never outlive their records.
#include "egs_common.cuh"

#define BL_ROWB 264u        // bytes of one parked row: w of the 32 pixels | u of the 32 pixels | 8 B pad (bank skew)

namespace {
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float ex2_approx_l(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx_l(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_add_v4_l(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32_l(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void sts32f_l(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float2 lds64f_l(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

struct LaneSmem {
    float4 rec[2][32 * 4];        // records of the chunk being walked / the chunk in flight (TMA destinations)
    uint32_t lm[2][32];           // their blend masks (this block's word)
    float2 park[32 * 33];         // parked w | u: row = hit of the chunk (264 B)
    unsigned long long bar[2];    // mbarriers of the two record buffers
    float4 ktab[16 * 4];          // per pixel pair: {gc0 e,o, gc1 e,o} {gc2 e,o, 10gn0 e,o} {10gn1 e,o, 10gn2 e,o} {gDn e,o, gD e,o}
};
} // namespace

#ifndef BL_MIN_CTAS
#define BL_MIN_CTAS 15
#endif
__global__ void __launch_bounds__(32, BL_MIN_CTAS)
k_render_backward_lane(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec,
                       ImgView im, BinView bn, long long cap, const float* __restrict__ gC,
                       const float* __restrict__ gN, const float* __restrict__ gDp, const float* __restrict__ gOp,
                       float* __restrict__ sg) {
    __shared__ __align__(16) LaneSmem S;

    const int tile = blockIdx.x >> 3, blk = blockIdx.x & 7;
    const long long start = im.tile_offset[tile];
    long long end = im.tile_offset[tile + 1];
    if (end > cap) end = cap;
    const int n = (int)(end - start);
    if (n <= 0) return;
    const int top = (int)min(im.hit_count[blockIdx.x], (uint32_t)n);   // entries of this block's hit list
    if (top <= 0) return;   // nothing was blended into this block

    const int tx = tile % gx, ty = tile / gx;
    const int lane = threadIdx.x;
    const int bx = tx * EGS_TILE + (blk & 1) * 8, by = ty * EGS_TILE + (blk >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)W * py + px;
    const float pxf = (float)px, pyf = (float)py;

    float T_final = 0.f, D_final = 0.f;
    float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gn0 = 0.f, gn1 = 0.f, gn2 = 0.f, gD = 0.f, gO = 0.f;
    if (inside) {
        T_final = im.final_T[pix];
        D_final = im.final_D[pix];
        // a pixel that blended nothing (the forward left T at its clamp) takes no part in any sum: drop its upstream
        // gradients so that a NaN / Inf there (a loss dividing by the rendered opacity, say) cannot reach a splat as
        // 0 * NaN.  The reference never visits such a pixel (backward.cu:540-560).
        if (T_final < 0.999999f) {
            gc0 = __ldg(gC + pix); gc1 = __ldg(gC + HW + pix); gc2 = __ldg(gC + 2 * HW + pix);
            gn0 = __ldg(gN + pix); gn1 = __ldg(gN + HW + pix); gn2 = __ldg(gN + 2 * HW + pix);
            gD = __ldg(gDp + pix);
            gO = __ldg(gOp + pix);
        }
    }

    const uint2* __restrict__ hseg = bn.hits + 8 * (size_t)start + (size_t)blk * (size_t)n;
    const int nc = (top + 31) >> 5;
    // chunk c covers hit-list positions top-1-32c ... top-32(c+1) (back to front); lane l holds position top-1-(32c+l)
    auto load_chunk = [&](int c, uint32_t& m, uint32_t& idv) {
        const int pos = top - 1 - (32 * c + lane);
        m = 0u;
        idv = 0u;
        if (pos >= 0) {
            const uint2 e = __ldg(hseg + pos);
            idv = e.x;
            m = e.y;
        }
    };
    const uint32_t rec_smem = smem_addr(S.rec);
    const uint32_t lm_smem = smem_addr(S.lm);
    const uint32_t bar_smem = smem_addr(S.bar);
    if (lane == 0) {
        mbar_init(bar_smem, 1u);
        mbar_init(bar_smem + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // every entry of a hit list has a non-empty mask, so a chunk is dense: slot = lane.  Chunk c lands in buffer c & 1.
    auto stage_chunk = [&](int c, uint32_t m, uint32_t idv) {
        const uint32_t buf = (uint32_t)(c & 1);
        const uint32_t bar = bar_smem + 8u * buf;
        if (lane == 0) mbar_expect_tx(bar, 64u * (uint32_t)min(32, top - 32 * c));
        if (m != 0u) {
            const uint32_t slot = buf * 32u + (uint32_t)lane;
            sts32(lm_smem + 4u * slot, m);
            bulk_copy_g2s(rec_smem + 64u * slot, rec + idv, 64u, bar);
        }
    };

    uint32_t m_nxt, id_nxt, id_cur;
    load_chunk(0, m_nxt, id_nxt);
    stage_chunk(0, m_nxt, id_nxt);
    id_cur = id_nxt;
    load_chunk(1, m_nxt, id_nxt);

    const float one_m_Tf = 1.f - T_final;
    const float gDn = gD / one_m_Tf;
    const float bg_dot = __ldg(bg) * gc0 + __ldg(bg + 1) * gc1 + __ldg(bg + 2) * gc2;
    const float K0 = gD * D_final / one_m_Tf / one_m_Tf * -T_final + T_final * (gO - bg_dot);
    const float kx = 2.f * 0.5f * (float)W, ky = 2.f * 0.5f * (float)H;

    // weight table of the block, pair-interleaved: pair j = lanes 2j (even) and 2j+1 (odd), horizontally adjacent pixels
    {
        const uint32_t kb = smem_addr(S.ktab) + 64u * (uint32_t)(lane >> 1) + 4u * (uint32_t)(lane & 1);
        sts32f_l(kb, gc0);            sts32f_l(kb + 8u, gc1);
        sts32f_l(kb + 16u, gc2);      sts32f_l(kb + 24u, gn0 * 10.f);     // x10: backward.cu:604
        sts32f_l(kb + 32u, gn1 * 10.f); sts32f_l(kb + 40u, gn2 * 10.f);
        sts32f_l(kb + 48u, gDn);      sts32f_l(kb + 56u, gD);
    }

    float T = T_final;
    float sigma = 0.f;
    const uint32_t park_base = smem_addr(S.park);
    const uint32_t ktab_base = smem_addr(S.ktab);
    const uint32_t park_lane = park_base + 4u * (uint32_t)lane;   // this pixel's w slot in a row; its u slot is 128 B further
    const uint32_t lanebit = 1u << lane;
    const float fbx = (float)bx, fby = (float)by;
    constexpr float LOG2E = 1.4426950408889634f;

    // phase 1 for one (splat, block) pair, branch-free.  A pixel that did not blend the splat gets G = 0, hence
    // alpha = w = u = 0, 1/(1-alpha) = 1 and an untouched T / sigma.
    auto pair_math = [&](uint32_t mask, const float4& q0, const float4& q1, const float4& q2, const float4& q3,
                         uint32_t prow) {
        const bool act = (mask & lanebit) != 0u;
        const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
        // -0.5 (cxx dx^2 + cyy dy^2) - cxy dx dy, in log2 units
        const float e = fmaf(q1.y, dy + dy, q1.x * dx);
        const float dist = fmaf(dx, e, (q1.z * dy) * dy);
        float G = ex2_approx_l(dist * (-0.5f * LOG2E));
        G = act ? G : 0.f;
        const float alpha = fminf(0.99f, q0.w * G);
        const float ra = rcp_approx_l(1.f - alpha);
        T = T * ra;                               // transmittance in front of this splat (ra == 1 for a non-blender)
        const float w = alpha * T;
        const float d_cur = q1.w - (dx * q2.x + dy * q2.y);
        float kappa = q2.z * gc0;
        kappa = fmaf(q2.w, gc1, kappa); kappa = fmaf(q3.x, gc2, kappa);
        kappa = fmaf(q3.y, gn0, kappa); kappa = fmaf(q3.z, gn1, kappa); kappa = fmaf(q3.w, gn2, kappa);
        kappa = fmaf(d_cur, gDn, kappa);
        const float dL_dalpha = fmaf(T, kappa, ra * (K0 - sigma));
        sigma = fmaf(w, kappa, sigma);
        sts32f_l(prow, w);
        sts32f_l(prow + 128u, G * dL_dalpha);     // u; dd = u * (-opacity / 2) is applied to the sums
    };

    for (int c = 0; c < nc; c++) {
        const uint32_t buf = (uint32_t)(c & 1);
        const int cnt = min(32, top - 32 * c);
        // phase 2 of chunk c-1 is done with the parked rows and with the other record buffer (generic-proxy reads):
        // order them before the async-proxy writes of the copies issued next
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // chunk c+1 -> the other buffer, chunk c+2 -> registers
        if (c + 1 < nc) stage_chunk(c + 1, m_nxt, id_nxt);
        const uint32_t id_mine = id_cur;      // surfel of hit `lane` of this chunk
        id_cur = id_nxt;
        load_chunk(c + 2, m_nxt, id_nxt);
        mbar_wait(bar_smem + 8u * buf, (uint32_t)((c >> 1) & 1));   // chunk c has landed
        __syncwarp();                                               // ... and everybody's mask words are visible

        const uint32_t rec_base = rec_smem + 2048u * buf;
        const uint32_t lm_base = lm_smem + 128u * buf;
        {
            uint32_t rad = rec_base, la = lm_base, prow = park_lane;
            int i = 0;
#pragma unroll 1
            for (; i + 2 <= cnt; i += 2) {   // two splats per iteration: both sets of loads are issued up front
                const uint32_t mA_ = lds32(la), mB_ = lds32(la + 4u);
                const float4 a0 = lds128(rad), a1 = lds128(rad + 16u), a2 = lds128(rad + 32u), a3 = lds128(rad + 48u);
                const float4 b0 = lds128(rad + 64u), b1 = lds128(rad + 80u), b2 = lds128(rad + 96u), b3 = lds128(rad + 112u);
                pair_math(mA_, a0, a1, a2, a3, prow);
                pair_math(mB_, b0, b1, b2, b3, prow + BL_ROWB);
                rad += 128u; la += 8u; prow += 2u * BL_ROWB;
            }
            if (i < cnt) {
                const uint32_t mA_ = lds32(la);
                const float4 a0 = lds128(rad), a1 = lds128(rad + 16u), a2 = lds128(rad + 32u), a3 = lds128(rad + 48u);
                pair_math(mA_, a0, a1, a2, a3, prow);
            }
        }
        __syncwarp();

        // phase 2: lane = hit `lane` of the chunk
        if (lane < cnt) {
            const uint32_t myrec = rec_base + 64u * (uint32_t)lane;
            const float4 q0 = lds128(myrec);
            const float xr = __fsub_rn(q0.x, fbx), yr = __fsub_rn(q0.y, fby);   // exact; dx = xr - lx equals phase 1's dx
            const uint32_t row = park_base + BL_ROWB * (uint32_t)lane;
            float2 s_ux = bc2(0.f), s_uy = s_ux, s_gd = s_ux, s_xx = s_ux, s_xy = s_ux, s_yy = s_ux, s_u = s_ux;
            float2 s_c0 = s_ux, s_c1 = s_ux, s_c2 = s_ux, s_n0 = s_ux, s_n1 = s_ux, s_n2 = s_ux, s_d = s_ux;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const float2 w2 = lds64f_l(row + 8u * (uint32_t)j);                // w even, w odd
                const float2 u2 = lds64f_l(row + 128u + 8u * (uint32_t)j);         // u even, u odd
                const float4 k0 = lds128(ktab_base + 64u * (uint32_t)j);            // broadcasts
                const float4 k1 = lds128(ktab_base + 64u * (uint32_t)j + 16u);
                const float4 k2 = lds128(ktab_base + 64u * (uint32_t)j + 32u);
                const float4 k3 = lds128(ktab_base + 64u * (uint32_t)j + 48u);
                const float lx = (float)(2 * (j & 3)), ly = (float)(j >> 2);
                const float2 dx2 = __fadd2_rn(bc2(xr), make_float2(-lx, -lx - 1.f));
                const float dy = yr - ly;
                const float2 udx = __fmul2_rn(u2, dx2), udy = __fmul2_rn(u2, bc2(dy));
                s_ux = __fadd2_rn(s_ux, udx);
                s_uy = __fadd2_rn(s_uy, udy);
                s_xx = __ffma2_rn(udx, dx2, s_xx);
                s_xy = __ffma2_rn(udx, bc2(dy), s_xy);         // conic xy (not doubled, like the reference)
                s_yy = __ffma2_rn(udy, bc2(dy), s_yy);
                s_u = __fadd2_rn(s_u, u2);
                // 1 where the pixel blended the splat (w >= alpha_min * T_min = 4e-7), else 0
                const float2 act = make_float2(__saturatef(w2.x * 1e30f), __saturatef(w2.y * 1e30f));
                s_gd = __ffma2_rn(make_float2(k3.z, k3.w), act, s_gd);
                s_c0 = __ffma2_rn(make_float2(k0.x, k0.y), w2, s_c0);
                s_c1 = __ffma2_rn(make_float2(k0.z, k0.w), w2, s_c1);
                s_c2 = __ffma2_rn(make_float2(k1.x, k1.y), w2, s_c2);
                s_n0 = __ffma2_rn(make_float2(k1.z, k1.w), w2, s_n0);
                s_n1 = __ffma2_rn(make_float2(k2.x, k2.y), w2, s_n1);
                s_n2 = __ffma2_rn(make_float2(k2.z, k2.w), w2, s_n2);
                s_d = __ffma2_rn(make_float2(k3.x, k3.y), w2, s_d);
            }
            const float4 q1 = lds128(myrec + 16u);
            const float4 q2 = lds128(myrec + 32u);
            const float hf = q0.w * -0.5f;                       // dL/d(dist) = u * opacity * (-1/2)
            const float a0 = (s_ux.x + s_ux.y) * hf, a1 = (s_uy.x + s_uy.y) * hf, a2 = s_gd.x + s_gd.y;
            const float v0 = kx * (q1.x * a0 + q1.y * a1) - q2.x * a2;   // backward.cu:648-660
            const float v1 = ky * (q1.z * a1 + q1.y * a0) - q2.y * a2;
            float* dst = sg + (size_t)EGS_SCREEN_GRAD_STRIDE * id_mine;
            red_add_v4_l(dst, v0, v1, (s_xx.x + s_xx.y) * hf, (s_xy.x + s_xy.y) * hf);
            red_add_v4_l(dst + 4, (s_yy.x + s_yy.y) * hf, s_u.x + s_u.y, s_c0.x + s_c0.y, s_c1.x + s_c1.y);
            red_add_v4_l(dst + 8, s_c2.x + s_c2.y, s_n0.x + s_n0.y, s_n1.x + s_n1.y, s_n2.x + s_n2.y);
            red_add_f32_l(dst + 12, s_d.x + s_d.y);
        }
    }
}

cudaError_t launch_render_backward_lane(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                        const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                        cudaStream_t s) {
    const int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
    cudaError_t e = cudaFuncSetAttribute(k_render_backward_lane, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k_render_backward_lane<<<gx * gy * 8, 32, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap, gC, gN, gD, gO, sg);
    return cudaGetLastError();
}
