// egs_render_bwd_gather.cu -- reverse compositing walk with a transposed ("gather") cross-pixel reduction.
//
// Same per-pixel arithmetic as egs_render_bwd.cu (see there for the algebra and the reference quirks it keeps).
// The shuffle butterfly there costs ~76 issue slots per (warp, splat); here the reduction is turned around:
//   phase 1 (lane = pixel):  for each splat that touches the warp's 8x4 block, every lane computes the three
//            per-pair scalars w = alpha*T, dd = dL/d(dist), u = G*dL/dalpha (+ its activity flag) and parks them as
//            one float4 in shared memory;
//   phase 2 (lane = splat x pixel-quarter), every 8 parked splats: lane (h, p) walks 8 of the 32 pixels of splat h,
//            multiplying the parked scalars with the pixels' constant weights (pixel gradients, pixel position) from
//            a per-warp table and accumulating the 14 sums in registers -- plain FP32 FMAs, all 32 lanes busy, no
//            selects.  Two shuffle steps merge the four quarters, and each of the four lanes of a splat issues one
//            16-byte vector reduction (red.global.add.v4.f32) of its quad of G[surfel][16].
//            Splats still parked when a staging buffer is about to be recycled have the three record words phase 2
//            needs copied into the padding units of their row, so phase 2 always runs on a full set of 8 (except
//            once, at the end of the tile) instead of being flushed half-empty after every batch.
// ~35 issue slots per (warp, splat) for the reduction, no CTA-wide combine pass and no per-batch barriers for it.
// Record staging is double-buffered with cp.async (LDGSTS): while a batch is walked, the next batch's 64-byte
// records and blend masks stream into the other buffer and the surfel ids of the batch after that are already in
// registers, so a tile pays the dependent-gather latency once instead of once per batch.
#include "egs_common.cuh"

#ifndef GB_BATCH
#define GB_BATCH 128   // A/B on C3: 32 -> 1.134 ms, 48 -> 1.061, 64 -> 1.056, 128 -> 1.037 (fewer CTA barriers)
#endif
#ifndef GB_MINCTAS
#define GB_MINCTAS 3
#endif
#define GB_WARPS (EGS_TILE_THREADS / 32)
#define GB_PEND 8          // splats parked per warp before a phase-2 pass
#define GB_ROW 36          // float4 units per parked splat row (32 pixels + padding: conflict-free both ways)

namespace {
__device__ __forceinline__ float conic_power_g(float cxx, float cxy, float cyy, float dx, float dy) {
    const float q = __fmaf_rn(__fmul_rn(cxx, dx), dx, __fmul_rn(__fmul_rn(cyy, dy), dy));
    const float dist = __fmaf_rn(__fmul_rn(__fmul_rn(2.f, cxy), dx), dy, q);
    return __fmul_rn(-0.5f, dist);
}
__device__ __forceinline__ float ex2_approx_g(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx_g(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_add_v4_g(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct GatherSmem {
    float4 rec[2][GB_BATCH * 4];
    uint32_t lm[2][GB_BATCH * 8];
    float4 pair[GB_WARPS][GB_PEND * GB_ROW];
    float4 ktab[GB_WARPS][32 * 2];
    int top;
};
} // namespace

__global__ void __launch_bounds__(EGS_TILE_THREADS, GB_MINCTAS)
k_render_backward_gather(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec,
                         ImgView im, BinView bn, long long cap, const float* __restrict__ gC,
                         const float* __restrict__ gN, const float* __restrict__ gDp, const float* __restrict__ gOp,
                         float* __restrict__ sg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GatherSmem& S = *reinterpret_cast<GatherSmem*>(smem_raw);

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

    if (threadIdx.x == 0) S.top = 0;
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
    if (lane == 0 && warp_last > 0) atomicMax(&S.top, warp_last);

    const float one_m_Tf = 1.f - T_final;
    const float gDn = gD / one_m_Tf;
    const float bg_dot = __ldg(bg) * gc0 + __ldg(bg + 1) * gc1 + __ldg(bg + 2) * gc2;
    const float K0 = gD * D_final / one_m_Tf / one_m_Tf * -T_final + T_final * (gO - bg_dot);
    const float kx = 2.f * 0.5f * (float)W, ky = 2.f * 0.5f * (float)H;

    // per-warp table of the pixels' constant weights, read (broadcast) in phase 2
    S.ktab[warp][lane * 2 + 0] = make_float4(gc0, gc1, gc2, gn0 * 10.f);       // x10: backward.cu:604
    S.ktab[warp][lane * 2 + 1] = make_float4(gn1 * 10.f, gn2 * 10.f, gDn, gD);
    __syncthreads();
    const int top0 = S.top;

    float T = T_final;
    float sigma = 0.f;
    uint32_t rec_base = smem_addr(S.rec[0]);
    const uint32_t* cur_lm = S.lm[0];
    const uint32_t pair_base = smem_addr(S.pair[warp]);
    const uint32_t ktab_base = smem_addr(S.ktab[warp]);
    const int h = lane >> 2, p = lane & 3;
    int npend = 0;
    uint32_t myrec = 0;   // shared address of the record (q0..q2; surfel id in q0.z) of parked splat h
    // phase-2 pixel coordinates: lane (h, p) visits pixels k = 4 i + p, i.e. block column p + 4 (i & 1), row i >> 1
    const float fpx0 = (float)(bx + p), fpy0 = (float)by;

    // phase 2 for the `np` parked splats of this warp
    auto reduce_pending = [&](int np) {
        __syncwarp();
        float a[14];
#pragma unroll
        for (int i = 0; i < 14; i++) a[i] = 0.f;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h < np) {
            q0 = lds128(myrec);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = 4 * i + p; // pixel (lane index of phase 1)
                const float4 pr = lds128(pair_base + 16u * (uint32_t)(h * GB_ROW + k)); // w, dd, u, act
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
            a[i] += __shfl_xor_sync(0xffffffffu, a[i], 1);
            a[i] += __shfl_xor_sync(0xffffffffu, a[i], 2);
        }
        if (h < np) {
            float* dst = sg + (size_t)EGS_SCREEN_GRAD_STRIDE * __float_as_uint(q0.z) + 4 * p;
            if (p == 0) {
                const float4 q1 = lds128(myrec + 16u);
                const float4 q2 = lds128(myrec + 32u);
                const float v0 = kx * (q1.x * a[0] + q1.y * a[1]) - q2.x * a[2];   // backward.cu:648-660
                const float v1 = ky * (q1.z * a[1] + q1.y * a[0]) - q2.y * a[2];
                red_add_v4_g(dst, v0, v1, a[3], a[4]);
            } else if (p == 1) {
                red_add_v4_g(dst, a[5], a[6], a[7], a[8]);
            } else if (p == 2) {
                red_add_v4_g(dst, a[9], a[10], a[11], a[12]);
            } else {
                red_add_v4_g(dst, a[13], 0.f, 0.f, 0.f);
            }
        }
        __syncwarp();
    };

    // ---- double-buffered staging.  Batch b covers list positions [top_b - m_b, top_b), top_b = top0 - b * GB_BATCH;
    // entry t of a batch is position top_b - 1 - t (back to front).
    const int tid = threadIdx.x;
    auto stage = [&](int buf, int top_b, uint32_t idv) {   // threads tid < GB_BATCH
        if (tid < min(GB_BATCH, top_b)) {
            const float4* src = reinterpret_cast<const float4*>(rec + idv);
#pragma unroll
            for (int q = 0; q < 4; q++) cp_async16(&S.rec[buf][tid * 4 + q], src + q);
            const uint4* lm = reinterpret_cast<const uint4*>(bn.lane_masks + 8 * (size_t)(start + top_b - 1 - tid));
            cp_async16(&S.lm[buf][tid * 8], lm);
            cp_async16(&S.lm[buf][tid * 8 + 4], lm + 1);
        } else {
            reinterpret_cast<uint4*>(S.lm[buf])[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(S.lm[buf])[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
    };
    const int nb = (top0 + GB_BATCH - 1) / GB_BATCH;
    uint32_t next_id = 0, staged_id = 0;
    if (tid < GB_BATCH && nb > 0) {
        if (tid < min(GB_BATCH, top0)) next_id = __ldg(plist + (top0 - 1 - tid));
        stage(0, top0, next_id);
        staged_id = next_id;
        const int top1 = top0 - GB_BATCH;
        if (nb > 1 && tid < min(GB_BATCH, top1)) next_id = __ldg(plist + (top1 - 1 - tid));
    }
    cp_async_commit();

    for (int b = 0; b < nb; b++) {
        const int buf = b & 1;
        cp_async_wait_all();
        // the extent word of the record (forward-only) is replaced by the surfel id: phase 2 finds it there
        if (tid < GB_BATCH) reinterpret_cast<uint32_t*>(&S.rec[buf][tid * 4])[2] = staged_id;
        __syncthreads(); // batch b has landed; every warp is done with batch b-1 (whose buffer is refilled next)
        if (tid < GB_BATCH) {
            const int top_n = top0 - (b + 1) * GB_BATCH;
            if (b + 1 < nb) stage(buf ^ 1, top_n, next_id);
            staged_id = next_id;
            const int top_nn = top_n - GB_BATCH;
            if (b + 2 < nb && tid < min(GB_BATCH, top_nn)) next_id = __ldg(plist + (top_nn - 1 - tid));
        }
        cp_async_commit();
        rec_base = smem_addr(S.rec[buf]);
        cur_lm = S.lm[buf];

#pragma unroll 1
        for (int c = 0; c < (GB_BATCH + 31) / 32; c++) {
            const int jl = c * 32 + lane;
            unsigned hits = __ballot_sync(0xffffffffu, (GB_BATCH % 32 == 0 || jl < GB_BATCH) && cur_lm[8 * jl + warp] != 0u);
            while (hits) {
                const int j = c * 32 + __ffs(hits) - 1; // batch entry j = list position top-1-j (back to front)
                hits &= hits - 1;
                const bool act = (cur_lm[8 * j + warp] >> lane) & 1u;
                const uint32_t rad = rec_base + 64u * (uint32_t)j;
                const float4 q0 = lds128(rad);
                const float4 q1 = lds128(rad + 16u);
                const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
                const float power = conic_power_g(q1.x, q1.y, q1.z, dx, dy);
                const float G = ex2_approx_g(power * 1.4426950408889634f);
                const float alpha = fminf(0.99f, q0.w * G);
                float w = 0.f, dd = 0.f, u = 0.f;
                if (act) {
                    const float4 q2 = lds128(rad + 32u), q3 = lds128(rad + 48u);
                    const float ra = rcp_approx_g(1.f - alpha);
                    T = T * ra;                       // transmittance in front of this splat
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
                sts128(pair_base + 16u * (uint32_t)(npend * GB_ROW + lane), w, dd, u, act ? 1.f : 0.f);
                if (h == npend) myrec = rad;
                if (++npend == GB_PEND) { reduce_pending(GB_PEND); npend = 0; }
            }
        }
        // Splats still parked point into this batch's staging buffer, which is refilled during the batch after
        // next: move their q0..q2 into the padding units 32..34 of their row (lane p copies unit p).
        if (h < npend) {
            const uint32_t keep = pair_base + 16u * (uint32_t)(h * GB_ROW + 32);
            if (p < 3) {
                const float4 v = lds128(myrec + 16u * (uint32_t)p);
                sts128(keep + 16u * (uint32_t)p, v.x, v.y, v.z, v.w);
            }
            myrec = keep;
        }
        __syncwarp();
    }
    if (npend) reduce_pending(npend);
}

cudaError_t launch_render_backward_gather(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                          const float* gC, const float* gN, const float* gD, const float* gO, float* sg,
                                          cudaStream_t s) {
    // per-device attribute, cheap to set: no static state, so the library stays re-entrant across devices
    cudaError_t e = cudaFuncSetAttribute(k_render_backward_gather, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(GatherSmem));
    if (e != cudaSuccess) return e;
    const int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
    k_render_backward_gather<<<gx * gy, EGS_TILE_THREADS, sizeof(GatherSmem), s>>>(f.width, f.height, gx, f.bg, g.rec, im,
                                                                                  bn, cap, gC, gN, gD, gO, sg);
    return cudaGetLastError();
}
