// egs_render_bwd_warp.cu -- reverse compositing walk, one WARP (one 8x4 pixel block) per CTA.
//
// Same arithmetic as egs_render_bwd_gather.cu (phase 1: lane = pixel, parks w / dd / u per pair; phase 2:
// lane = splat x pixel-quarter, gathers the parked pairs into the 14 sums and issues red.global.add.v4.f32), see
// there and egs_render_bwd.cu for the algebra and the reference quirks that are kept.  What changes is the
// scheduling unit.  With one CTA per 16x16 tile the eight warps of a tile meet at a barrier once per staged
// batch although their work differs (a splat touches 2.7 of the 8 blocks on average, unevenly): ncu showed
// 36 % of the stall samples on that barrier.  Here every 8x4 block is its own 32-thread CTA:
//   * no CTA-wide barriers at all (only __syncwarp), no tile tail that keeps seven finished warps resident;
//   * a warp stages ONLY the splats that blended into its block: the forward left it a compacted, depth-ordered
//     hit list {surfel id, pixel mask} (egs_render_fwd.cu, HITLIST), so a chunk is 32 coalesced 8-byte entries, all
//     of them real hits, whose 64-byte records are cp.async-copied into the warp's own double buffer (the next
//     chunk's records and the entries of the chunk after that are in flight while the current chunk is walked);
//   * the walk covers exactly the block's own contributors, back to front.
// Shared memory is 10 KB per warp, so 20 warps are resident per SM, all of them runnable.
#include "egs_common.cuh"

#define WB_PEND 8          // splats parked before a phase-2 pass
#define WB_ROW 36          // float4 units per parked row (32 pixels + 4 padding units, reused to carry records)

namespace {
__device__ __forceinline__ float conic_power_w(float cxx, float cxy, float cyy, float dx, float dy) {
    const float q = __fmaf_rn(__fmul_rn(cxx, dx), dx, __fmul_rn(__fmul_rn(cyy, dy), dy));
    const float dist = __fmaf_rn(__fmul_rn(__fmul_rn(2.f, cxy), dx), dy, q);
    return __fmul_rn(-0.5f, dist);
}
__device__ __forceinline__ float ex2_approx_w(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx_w(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_add_v4_w(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128_w(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cp_async16_w(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_w() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1_w() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct WarpSmem {
    float4 rec[2][32 * 4];            // compacted records of the chunk being walked / the chunk in flight
    uint32_t lm[2][32];               // their blend masks (this block's word)
    uint32_t id[2][32];               // their surfel ids
    float4 pair[WB_PEND * WB_ROW];    // parked pairs
    float4 ktab[32 * 2];              // the pixels' constant weights
};
} // namespace

// Resident CTAs per SM the register allocation aims for.  Shared memory (10 KB + 1 KB per CTA) caps residency at 20;
// asking for 21 makes ptxas fit the kernel in 80 registers without spills, which measured 0.946 ms against 0.954 ms
// at 96 registers; 16 resident warps (122 registers): 0.968 ms -- occupancy is not the lever here.
#ifndef WB_MIN_CTAS
#define WB_MIN_CTAS 21
#endif
__global__ void __launch_bounds__(32, WB_MIN_CTAS)
k_render_backward_warp(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec,
                       ImgView im, BinView bn, long long cap, const float* __restrict__ gC,
                       const float* __restrict__ gN, const float* __restrict__ gDp, const float* __restrict__ gOp,
                       float* __restrict__ sg) {
    __shared__ __align__(16) WarpSmem S;
    const unsigned full = 0xffffffffu;

    const int tile = blockIdx.x >> 3, blk = blockIdx.x & 7;
    const long long start = im.tile_offset[tile];
    long long end = im.tile_offset[tile + 1];
    if (end > cap) end = cap;
    const int n = (int)(end - start);
    if (n <= 0) return;

    const int tx = tile % gx, ty = tile / gx;
    const int lane = threadIdx.x;
    const int bx = tx * EGS_TILE + (blk & 1) * 8, by = ty * EGS_TILE + (blk >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)W * py + px;
    const float pxf = (float)px, pyf = (float)py;

    const int top = (int)min(im.hit_count[blockIdx.x], (uint32_t)n);   // entries of this block's hit list
    if (top <= 0) return;   // nothing was blended into this block

    float T_final = 0.f, D_final = 0.f;
    float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gn0 = 0.f, gn1 = 0.f, gn2 = 0.f, gD = 0.f, gO = 0.f;
    if (inside) {
        T_final = im.final_T[pix];
        D_final = im.final_D[pix];
        gc0 = __ldg(gC + pix); gc1 = __ldg(gC + HW + pix); gc2 = __ldg(gC + 2 * HW + pix);
        gn0 = __ldg(gN + pix); gn1 = __ldg(gN + HW + pix); gn2 = __ldg(gN + 2 * HW + pix);
        gD = __ldg(gDp + pix);
        gO = __ldg(gOp + pix);
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
    const uint32_t id_smem = smem_addr(S.id);
    // compact the chunk's hits into staging buffer `buf`; returns their count
    // (every entry of a hit list has a non-empty mask, so a chunk is dense: slot = lane)
    auto stage_chunk = [&](int c, int buf, uint32_t m, uint32_t idv) -> int {
        if (m != 0u) {
            const uint32_t slot = (uint32_t)buf * 32u + (uint32_t)lane;
            sts32(lm_smem + 4u * slot, m);
            sts32(id_smem + 4u * slot, idv);
            const float4* src = reinterpret_cast<const float4*>(rec + idv);
            const uint32_t dst = rec_smem + 64u * slot;
#pragma unroll
            for (int q = 0; q < 4; q++) cp_async16_w(dst + 16u * q, src + q);
        }
        return min(32, top - 32 * c);
    };

    uint32_t mA, idA;
    load_chunk(0, mA, idA);
    int cnt_cur = stage_chunk(0, 0, mA, idA);
    cp_async_commit_w();
    load_chunk(1, mA, idA);

    const float one_m_Tf = 1.f - T_final;
    const float gDn = gD / one_m_Tf;
    const float bg_dot = __ldg(bg) * gc0 + __ldg(bg + 1) * gc1 + __ldg(bg + 2) * gc2;
    const float K0 = gD * D_final / one_m_Tf / one_m_Tf * -T_final + T_final * (gO - bg_dot);
    const float kx = 2.f * 0.5f * (float)W, ky = 2.f * 0.5f * (float)H;

    S.ktab[lane * 2 + 0] = make_float4(gc0, gc1, gc2, gn0 * 10.f);       // x10: backward.cu:604
    S.ktab[lane * 2 + 1] = make_float4(gn1 * 10.f, gn2 * 10.f, gDn, gD);

    float T = T_final;
    float sigma = 0.f;
    const uint32_t pair_base = smem_addr(S.pair);
    const uint32_t ktab_base = smem_addr(S.ktab);
    const int h = lane >> 2, p = lane & 3;
    const uint32_t lanebit = 1u << lane;
    const uint32_t keep = pair_base + 16u * (uint32_t)(h * WB_ROW + 32);   // carried record of parked row h
    int npend = 0;       // parked splats (rows 0 .. npend-1)
    int carried = 0;     // rows < carried have their record in `keep`; the others at row_rec + 64 * row
    uint32_t row_rec = 0;
    const float fpx0 = (float)(bx + p), fpy0 = (float)by;

    // phase 1 for one (splat, block) pair: branch-free.  Pixels that did not blend the splat get G = 0, hence
    // alpha = w = dd = u = 0 and an untouched sigma; only the transmittance update needs its own predicate.
    auto pair_math = [&](uint32_t mask, const float4& q0, const float4& q1, const float4& q2, const float4& q3,
                         uint32_t prow) {
        const bool act = (mask & lanebit) != 0u;
        const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
        const float power = conic_power_w(q1.x, q1.y, q1.z, dx, dy);
        float G = ex2_approx_w(power * 1.4426950408889634f);
        G = act ? G : 0.f;
        const float alpha = fminf(0.99f, q0.w * G);
        const float ra = rcp_approx_w(1.f - alpha);
        if (act) T = T * ra;                  // transmittance in front of this splat
        const float w = alpha * T;
        const float d_cur = q1.w - (dx * q2.x + dy * q2.y);
        float kappa = q2.z * gc0;
        kappa = fmaf(q2.w, gc1, kappa); kappa = fmaf(q3.x, gc2, kappa);
        kappa = fmaf(q3.y, gn0, kappa); kappa = fmaf(q3.z, gn1, kappa); kappa = fmaf(q3.w, gn2, kappa);
        kappa = fmaf(d_cur, gDn, kappa);
        const float dL_dalpha = fmaf(T, kappa, ra * (K0 - sigma));
        sigma = fmaf(w, kappa, sigma);
        const float u = G * dL_dalpha;
        const float dd = u * (q0.w * -0.5f);
        sts128_w(prow, w, dd, u, act ? 1.f : 0.f);
    };

    auto reduce_pending = [&](int np) {
        __syncwarp();
        float a[14];
#pragma unroll
        for (int i = 0; i < 14; i++) a[i] = 0.f;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t myrec = h < carried ? keep : row_rec + 64u * (uint32_t)h;   // q0..q2, surfel id in q0.z
        if (h < np) {
            q0 = lds128(myrec);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = 4 * i + p; // pixel (lane index of phase 1)
                const float4 pr = lds128(pair_base + 16u * (uint32_t)(h * WB_ROW + k)); // w, dd, u, act
                const float4 k0 = lds128(ktab_base + 16u * (uint32_t)(k * 2));
                const float4 k1 = lds128(ktab_base + 16u * (uint32_t)(k * 2 + 1));
                const float dx = __fsub_rn(q0.x, fpx0 + (float)(4 * (i & 1))), dy = __fsub_rn(q0.y, fpy0 + (float)(i >> 1));
                const float ddx = pr.y * dx, ddy = pr.y * dy;
                a[0] += ddx;                    // sum dd*dx
                a[1] += ddy;                    // sum dd*dy
                a[2] = fmaf(k1.w, pr.w, a[2]);  // sum gD over the active pixels
                a[3] = fmaf(ddx, dx, a[3]);     // conic xx
                a[4] = fmaf(ddx, dy, a[4]);     // conic xy (not doubled, like the reference)
                a[5] = fmaf(ddy, dy, a[5]);     // conic yy
                a[6] += pr.z;                   // opacity
                a[7] = fmaf(k0.x, pr.x, a[7]); a[8] = fmaf(k0.y, pr.x, a[8]); a[9] = fmaf(k0.z, pr.x, a[9]);
                a[10] = fmaf(k0.w, pr.x, a[10]); a[11] = fmaf(k1.x, pr.x, a[11]); a[12] = fmaf(k1.y, pr.x, a[12]);
                a[13] = fmaf(k1.z, pr.x, a[13]);
            }
        }
#pragma unroll
        for (int i = 0; i < 14; i++) {
            a[i] += __shfl_xor_sync(full, a[i], 1);
            a[i] += __shfl_xor_sync(full, a[i], 2);
        }
        if (h < np) {
            float* dst = sg + (size_t)EGS_SCREEN_GRAD_STRIDE * __float_as_uint(q0.z) + 4 * p;
            if (p == 0) {
                const float4 q1 = lds128(myrec + 16u);
                const float4 q2 = lds128(myrec + 32u);
                const float v0 = kx * (q1.x * a[0] + q1.y * a[1]) - q2.x * a[2];   // backward.cu:648-660
                const float v1 = ky * (q1.z * a[1] + q1.y * a[0]) - q2.y * a[2];
                red_add_v4_w(dst, v0, v1, a[3], a[4]);
            } else if (p == 1) {
                red_add_v4_w(dst, a[5], a[6], a[7], a[8]);
            } else if (p == 2) {
                red_add_v4_w(dst, a[9], a[10], a[11], a[12]);
            } else {
                red_add_v4_w(dst, a[13], 0.f, 0.f, 0.f);
            }
        }
        __syncwarp();
    };

    for (int c = 0; c < nc; c++) {
        const uint32_t buf = (uint32_t)(c & 1);
        // chunk c+1 -> the other buffer (chunk c-1 has been walked and its parked splats moved out), chunk c+2 -> regs
        int cnt_next = 0;
        if (c + 1 < nc) cnt_next = stage_chunk(c + 1, (int)(buf ^ 1u), mA, idA);
        cp_async_commit_w();
        load_chunk(c + 2, mA, idA);
        cp_async_wait1_w();   // chunk c has landed (this thread's copies) ...
        __syncwarp();         // ... and everybody else's
        // the extent word of a record (forward-only) is replaced by the surfel id: phase 2 finds it there
        if (lane < cnt_cur) sts32(rec_smem + 64u * (buf * 32u + (uint32_t)lane) + 8u, lds32(id_smem + 4u * (buf * 32u + (uint32_t)lane)));
        __syncwarp();

        const uint32_t rec_base = rec_smem + 2048u * buf;
        const uint32_t lm_base = lm_smem + 128u * buf;
        int r = 0;
        while (r < cnt_cur) {
            // the next m hits fill the parked rows npend .. npend+m-1
            const int m = min(cnt_cur - r, WB_PEND - npend);
            uint32_t rad = rec_base + 64u * (uint32_t)r;
            uint32_t la = lm_base + 4u * (uint32_t)r;
            uint32_t prow = pair_base + 16u * (uint32_t)(npend * WB_ROW + lane);
            if (npend == carried) row_rec = rad - 64u * (uint32_t)npend;   // first uncarried row of this set
            int i = 0;
#pragma unroll 1
            for (; i + 2 <= m; i += 2) {   // two splats per iteration: both sets of loads are issued up front
                const uint32_t mA_ = lds32(la), mB_ = lds32(la + 4u);
                const float4 a0 = lds128(rad), a1 = lds128(rad + 16u), a2 = lds128(rad + 32u), a3 = lds128(rad + 48u);
                const float4 b0 = lds128(rad + 64u), b1 = lds128(rad + 80u), b2 = lds128(rad + 96u), b3 = lds128(rad + 112u);
                pair_math(mA_, a0, a1, a2, a3, prow);
                pair_math(mB_, b0, b1, b2, b3, prow + 16u * WB_ROW);
                rad += 128u; la += 8u; prow += 32u * WB_ROW;
            }
            if (i < m) {
                const uint32_t mA_ = lds32(la);
                const float4 a0 = lds128(rad), a1 = lds128(rad + 16u), a2 = lds128(rad + 32u), a3 = lds128(rad + 48u);
                pair_math(mA_, a0, a1, a2, a3, prow);
            }
            r += m;
            npend += m;
            if (npend == WB_PEND) { reduce_pending(WB_PEND); npend = 0; carried = 0; }
        }
        // Splats still parked point into this chunk's staging buffer, which is refilled next iteration: move their
        // q0..q2 into the padding units 32..34 of their row (lane p copies unit p).
        if (h >= carried && h < npend && p < 3) {
            const float4 v = lds128(row_rec + 64u * (uint32_t)h + 16u * (uint32_t)p);
            sts128_w(keep + 16u * (uint32_t)p, v.x, v.y, v.z, v.w);
        }
        carried = npend;
        __syncwarp();
        cnt_cur = cnt_next;
    }
    if (npend) reduce_pending(npend);
}

cudaError_t launch_render_backward_warp(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                        const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                        cudaStream_t s) {
    const int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
    cudaError_t e = cudaFuncSetAttribute(k_render_backward_warp, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k_render_backward_warp<<<gx * gy * 8, 32, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap, gC, gN, gD, gO, sg);
    return cudaGetLastError();
}
