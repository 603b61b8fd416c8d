// egs_render_fwd.cu -- front-to-back compositing of colour / normal / plane-corrected depth / opacity.
//
// Replaces renderCUDA<3> forward (DGS/cuda_rasterizer/forward.cu:306-497).  Differences in organisation, not in
// results:
//   * one CTA per tile of the WHOLE grid: tiles with an empty list write zeros (the reference leaves its
//     zero-initialised outputs untouched there), so no memset and no host-side tile compaction is needed;
//   * a warp owns an 8x4 pixel block (full 32-byte sectors on every image row it writes).  While a batch is
//     staged, the staging thread of each record computes which of the 8 blocks its alpha >= 1/255 ellipse box
//     can reach (one mask word); a warp then visits only its own hits (ballot + find-first-set).  The kernel is
//     issue-bound (ncu: 89 % issue-active, DRAM 2 %), so instructions per (tile, surfel) are what matters;
//   * splats are staged as packed 64-byte records (one gather per instance instead of eight);
//   * every warp appends {surfel id, 32-bit mask of the pixels that actually blended it} to the hit list of its
//     8x4 block (HITLIST, 8 B per hit, depth order): the backward walks exactly those pairs, never repeats the
//     alpha / transmittance tests and never scans list entries that missed its block.  The tile-wide backward
//     variants get the same information as 8 mask words per instance (lane_masks) instead.
// Per pixel the arithmetic order of the reference is kept: power, alpha = min(0.99, o*exp(power)), skip < 1/255,
// stop (without blending) when T(1-alpha) < 1e-4, w = alpha*T, fma accumulation, T clamp at 1-1e-6.
#include "egs_common.cuh"
#include <stdlib.h>

#define FWD_BATCH 256

__device__ __forceinline__ float conic_power(float cxx, float cxy, float cyy, float dx, float dy) {
    // (cxx*dx*dx + cyy*dy*dy) + 2*cxy*dx*dy, grouped as nvcc contracts the reference expression
    const float q = __fmaf_rn(__fmul_rn(cxx, dx), dx, __fmul_rn(__fmul_rn(cyy, dy), dy));
    const float dist = __fmaf_rn(__fmul_rn(__fmul_rn(2.f, cxy), dx), dy, q);
    return __fmul_rn(-0.5f, dist);
}

// Which of the tile's 8 warp blocks (2 columns x 4 rows of 8x4 pixels) can the record's extent box reach?
__device__ __forceinline__ uint32_t block_mask_f(float x, float y, uint32_t ext, float tile_x0, float tile_y0) {
    const float hx = (float)(ext & 0xffffu) * 0.125f + 3.5f, hy = (float)(ext >> 16) * 0.125f + 1.5f;
    const float rx = x - tile_x0, ry = y - tile_y0;
    const uint32_t xm = (fabsf(rx - 3.5f) <= hx ? 1u : 0u) | (fabsf(rx - 11.5f) <= hx ? 2u : 0u);
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (fabsf(ry - (1.5f + 4.f * r)) <= hy) m |= xm << (2 * r);
    return m;
}

// MODE 0: blend masks as 8 words per instance (tile-wide backward variants); 1: per-block hit lists (default);
// 2: forward-only render (EGS_FWD_NO_SAVE): nothing is saved for a backward.
// Resident CTAs per SM the register allocation aims for (measured at C3, stage emit + sort + forward): 5 -> 48
// registers, no spills, 0.7425 ms; 6 -> 40 registers with an 8-byte spill, 0.7512 ms; 7 / 8 -> 32 registers, 0.770 ms.
#ifndef FWD_MIN_CTAS
#define FWD_MIN_CTAS 5
#endif
template <int MODE>
__global__ void __launch_bounds__(EGS_TILE_THREADS, FWD_MIN_CTAS)
k_render_forward(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec, ImgView im,
                 BinView bn, long long cap, float* __restrict__ out_color, float* __restrict__ out_normal,
                 float* __restrict__ out_depth, float* __restrict__ out_opac) {
    __shared__ float4 s_rec[FWD_BATCH * 4];
    __shared__ uint32_t s_wm[FWD_BATCH];
    constexpr bool HITLIST = MODE == 1, SAVE = MODE != 2;
    __shared__ __align__(16) uint32_t s_lm[SAVE ? FWD_BATCH * 8 : 4];   // blend masks of the batch: [instance][block], HITLIST: [block][instance]

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = tx * EGS_TILE + (warp & 1) * 8, by = ty * EGS_TILE + (warp >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)W * py + px;

    const long long start = im.tile_offset[tile];
    long long end = im.tile_offset[tile + 1];
    if (end > cap) end = cap;
    const int n = (int)(end - start);
    if (n <= 0) {
        if (inside) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) { out_color[ch * HW + pix] = 0.f; out_normal[ch * HW + pix] = 0.f; }
            out_depth[pix] = 0.f;
            out_opac[pix] = 0.f;
            if (SAVE) {
                im.final_T[pix] = 0.f;   // saved state of never-composited tiles reads as zero (deterministic workspace)
                im.final_D[pix] = 0.f;
                im.n_contrib[pix] = 0u;
            }
        }
        if (HITLIST && threadIdx.x < 8) im.hit_count[8 * tile + threadIdx.x] = 0u;
        return;
    }
    const uint32_t* __restrict__ plist = bn.point_list + start;
    const float pxf = (float)px, pyf = (float)py;
    const float tile_x0 = (float)(tx * EGS_TILE), tile_y0 = (float)(ty * EGS_TILE);

    const uint32_t rec_base = smem_addr(s_rec);
    const uint32_t wm_lane = smem_addr(s_wm) + 4u * (uint32_t)lane;    // this lane's slot of a 32-entry chunk
    const uint32_t lm_warp = smem_addr(s_lm) + (HITLIST ? 4u * FWD_BATCH : 4u) * (uint32_t)warp;   // this warp's words
    constexpr uint32_t LM_STRIDE = HITLIST ? 4u : 32u;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f, D = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    // this block's hit list: entries 8*start + warp*n ... (at most n of them)
    uint2* __restrict__ hseg = bn.hits + 8 * (size_t)start + (size_t)warp * (size_t)n;
    uint32_t hcnt = 0u;

    for (int base = 0; base < n; base += FWD_BATCH) {
        // also the barrier that keeps the previous batch's records alive until every warp is done with them
        if (__syncthreads_count(done) == EGS_TILE_THREADS) break;
        const int m = min(FWD_BATCH, n - base);
        uint32_t wm = 0u;
        if ((int)threadIdx.x < m) {
            const uint32_t id = __ldg(plist + base + threadIdx.x);
            const float4* src = reinterpret_cast<const float4*>(rec + id);
            const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
            wm = block_mask_f(a.x, a.y, __float_as_uint(a.z), tile_x0, tile_y0);
            // the extent word has served its purpose: the staged copy carries the surfel id in its place
            s_rec[threadIdx.x * 4] = HITLIST ? make_float4(a.x, a.y, __uint_as_float(id), a.w) : a;
            s_rec[threadIdx.x * 4 + 1] = b;
            s_rec[threadIdx.x * 4 + 2] = c; s_rec[threadIdx.x * 4 + 3] = d;
        }
        s_wm[threadIdx.x] = wm;
        if (SAVE) {
            reinterpret_cast<uint4*>(s_lm)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(s_lm)[threadIdx.x + FWD_BATCH] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, done)) {
            const int chunks = (m + 31) >> 5;
            for (int c = 0; c < chunks; c++) {
                unsigned hits = __ballot_sync(0xffffffffu, (lds32(wm_lane + 128u * (uint32_t)c) >> warp) & 1u);
                while (hits) {
                    const int j = c * 32 + __ffs(hits) - 1;
                    hits &= hits - 1;
                    const uint32_t ra = rec_base + 64u * (uint32_t)j;
                    const float4 q0 = lds128(ra);
                    const float4 q1 = lds128(ra + 16u);
                    const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
                    const float power = conic_power(q1.x, q1.y, q1.z, dx, dy);
                    const float alpha = fminf(0.99f, __fmul_rn(q0.w, expf(power)));
                    const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
                    bool ok = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    if (ok && test_T < 0.0001f) { done = true; ok = false; }   // stops WITHOUT blending this one
                    if (SAVE) {
                        const unsigned bm = __ballot_sync(0xffffffffu, ok);
                        if (bm == 0u) continue;
                        sts32(lm_warp + LM_STRIDE * (uint32_t)j, bm);   // every lane stores the same word: one wavefront
                    }
                    if (ok) {
                        const float w = __fmul_rn(alpha, T);
                        const float4 q2 = lds128(ra + 32u), q3 = lds128(ra + 48u);
                        const float dj = q1.w - (dx * q2.x + dy * q2.y);
                        D = fmaf(dj, w, D);
                        C0 = fmaf(q2.z, w, C0); C1 = fmaf(q2.w, w, C1); C2 = fmaf(q3.x, w, C2);
                        N0 = fmaf(q3.y, w, N0); N1 = fmaf(q3.z, w, N1); N2 = fmaf(q3.w, w, N2);
                        T = test_T;
                        last = (uint32_t)(base + j + 1);
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
            if (HITLIST) {
                // append this batch's blended entries to the block's hit list, in list order: only this warp wrote
                // (and reads) its row of s_lm, so no CTA barrier is needed
                __syncwarp();
                for (int c = 0; c < chunks; c++) {
                    const uint32_t j = (uint32_t)(c * 32 + lane);
                    const uint32_t mk = lds32(lm_warp + 4u * j);
                    const unsigned hb = __ballot_sync(0xffffffffu, mk != 0u);
                    if (mk != 0u)
                        hseg[hcnt + (uint32_t)__popc(hb & ((1u << lane) - 1u))] = make_uint2(lds32(rec_base + 64u * j + 8u), mk);
                    hcnt += (uint32_t)__popc(hb);
                }
            }
        }
        if (MODE == 0) {
            __syncthreads();
            // publish the blend masks of this batch: 32 contiguous bytes per instance
            if ((int)threadIdx.x < m) {
                uint4* dst = reinterpret_cast<uint4*>(bn.lane_masks + 8 * (size_t)(start + base + threadIdx.x));
                dst[0] = reinterpret_cast<const uint4*>(s_lm)[2 * threadIdx.x];
                dst[1] = reinterpret_cast<const uint4*>(s_lm)[2 * threadIdx.x + 1];
            }
        }
    }
    if (HITLIST && lane == 0) im.hit_count[8 * tile + warp] = hcnt;
    if (inside) {
        T = fminf(0.999999f, T);
        if (SAVE) {
            im.final_T[pix] = T;
            im.final_D[pix] = D;
            im.n_contrib[pix] = last;
        }
        out_color[pix] = fmaf(T, __ldg(bg), C0);
        out_color[HW + pix] = fmaf(T, __ldg(bg + 1), C1);
        out_color[2 * HW + pix] = fmaf(T, __ldg(bg + 2), C2);
        out_normal[pix] = N0; out_normal[HW + pix] = N1; out_normal[2 * HW + pix] = N2;
        out_depth[pix] = D / (1.f - T);
        out_opac[pix] = 1.f - T;
    }
}

// ---- variant with TMA-staged record batches ---------------------------------------------------------------------
// Same arithmetic and outputs as k_render_forward<true>.  What changes is how a batch of records reaches shared
// memory: every staging thread issues ONE 64-byte bulk copy (cp.async.bulk, the TMA engine; SASS UBLKCP) that
// completes on an mbarrier, instead of 4 x LDG.128 + 4 x STS.128 through its registers, and the batches are
// double-buffered: the copies of batch b+1 (whose surfel ids were prefetched one batch earlier) are in flight while
// the warps walk batch b, so the gather latency is off the critical path.
// Measured on B200 at C3 (bench.py stage "render" = emit + sort + forward): default (LDG + STS staging, 6 CTAs/SM)
// 0.752 ms; this kernel with 256-record batches (2 x 16 KB, 5 CTAs/SM) 0.781 ms, with 128-record batches (6 CTAs/SM,
// twice the barriers) 0.807 ms.  The forward is instruction-issue bound (86 % issue-active), a 64-byte record is the
// smallest granule the copy engine moves, and the second buffer costs occupancy -- so the variant is correct
// (bit-identical, tests/test_parity_gpu.py) but NOT the default; it is selected with EGS_FWD_KERNEL=bulk.
#ifndef FWD_BULK
#define FWD_BULK 256
#endif

__global__ void __launch_bounds__(EGS_TILE_THREADS, FWD_BULK > 128 ? 5 : 6)
k_render_forward_bulk(int W, int H, int gx, const float* __restrict__ bg, const SplatRecord* __restrict__ rec,
                      ImgView im, BinView bn, long long cap, float* __restrict__ out_color,
                      float* __restrict__ out_normal, float* __restrict__ out_depth, float* __restrict__ out_opac) {
    __shared__ __align__(128) float4 s_rec[2][FWD_BULK * 4];
    __shared__ uint32_t s_wm[FWD_BULK];
    __shared__ __align__(16) uint32_t s_lm[FWD_BULK * 8];   // [block][instance]
    __shared__ __align__(8) unsigned long long s_bar[2];

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx = tx * EGS_TILE + (warp & 1) * 8, by = ty * EGS_TILE + (warp >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)W * py + px;

    const long long start = im.tile_offset[tile];
    long long end = im.tile_offset[tile + 1];
    if (end > cap) end = cap;
    const int n = (int)(end - start);
    if (n <= 0) {
        if (inside) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) { out_color[ch * HW + pix] = 0.f; out_normal[ch * HW + pix] = 0.f; }
            out_depth[pix] = 0.f;
            out_opac[pix] = 0.f;
            im.final_T[pix] = 0.f;
            im.final_D[pix] = 0.f;
            im.n_contrib[pix] = 0u;
        }
        if (tid < 8) im.hit_count[8 * tile + tid] = 0u;
        return;
    }
    const uint32_t* __restrict__ plist = bn.point_list + start;
    const float pxf = (float)px, pyf = (float)py;
    const float tile_x0 = (float)(tx * EGS_TILE), tile_y0 = (float)(ty * EGS_TILE);
    const uint32_t rec_smem = smem_addr(s_rec);
    const uint32_t bar_smem = smem_addr(s_bar);
    const uint32_t wm_lane = smem_addr(s_wm) + 4u * (uint32_t)lane;
    const uint32_t lm_warp = smem_addr(s_lm) + 4u * FWD_BULK * (uint32_t)warp;
    if (tid == 0) {
        mbar_init(bar_smem, 1u);
        mbar_init(bar_smem + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nb = (n + FWD_BULK - 1) / FWD_BULK;
    auto batch_len = [&](int b) { return min(FWD_BULK, n - b * FWD_BULK); };
    auto load_id = [&](int b) -> uint32_t { return (b < nb && tid < batch_len(b)) ? __ldg(plist + b * FWD_BULK + tid) : 0u; };
    auto issue = [&](int b, uint32_t id) {
        const int m = batch_len(b);
        const uint32_t bar = bar_smem + 8u * (uint32_t)(b & 1);
        if (tid == 0) mbar_expect_tx(bar, 64u * (uint32_t)m);
        if (tid < m) bulk_copy_g2s(rec_smem + 64u * FWD_BULK * (uint32_t)(b & 1) + 64u * (uint32_t)tid, rec + id, 64u, bar);
    };
    uint32_t id_cur = load_id(0);
    issue(0, id_cur);
    uint32_t id_nxt = load_id(1);

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f, D = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    uint2* __restrict__ hseg = bn.hits + 8 * (size_t)start + (size_t)warp * (size_t)n;
    uint32_t hcnt = 0u;

    for (int b = 0; b < nb; b++) {
        const uint32_t id_b1 = id_nxt;
        if (b + 1 < nb) {
            // buffer (b+1)&1 was released by the closing barrier of batch b-1; order our generic-proxy accesses to it
            // before the async-proxy writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(b + 1, id_b1);
            id_nxt = load_id(b + 2);
        }
        const int m = batch_len(b);
        const uint32_t rec_base = rec_smem + 64u * FWD_BULK * (uint32_t)(b & 1);
        mbar_wait(bar_smem + 8u * (uint32_t)(b & 1), (uint32_t)((b >> 1) & 1));
        if (tid < FWD_BULK) {
            uint32_t wm = 0u;
            if (tid < m) {
                const float4 a = lds128(rec_base + 64u * (uint32_t)tid);
                wm = block_mask_f(a.x, a.y, __float_as_uint(a.z), tile_x0, tile_y0);
                sts32(rec_base + 64u * (uint32_t)tid + 8u, id_cur);   // the extent word is replaced by the surfel id
            }
            s_wm[tid] = wm;
        }
#pragma unroll
        for (int z = tid; z < 2 * FWD_BULK; z += EGS_TILE_THREADS) reinterpret_cast<uint4*>(s_lm)[z] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (!__all_sync(0xffffffffu, done)) {
            const int chunks = (m + 31) >> 5;
            for (int c = 0; c < chunks; c++) {
                unsigned hits = __ballot_sync(0xffffffffu, (lds32(wm_lane + 128u * (uint32_t)c) >> warp) & 1u);
                while (hits) {
                    const int j = c * 32 + __ffs(hits) - 1;
                    hits &= hits - 1;
                    const uint32_t ra = rec_base + 64u * (uint32_t)j;
                    const float4 q0 = lds128(ra);
                    const float4 q1 = lds128(ra + 16u);
                    const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
                    const float power = conic_power(q1.x, q1.y, q1.z, dx, dy);
                    const float alpha = fminf(0.99f, __fmul_rn(q0.w, expf(power)));
                    const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
                    bool ok = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    if (ok && test_T < 0.0001f) { done = true; ok = false; }   // stops WITHOUT blending this one
                    const unsigned bm = __ballot_sync(0xffffffffu, ok);
                    if (bm == 0u) continue;
                    sts32(lm_warp + 4u * (uint32_t)j, bm);
                    if (ok) {
                        const float w = __fmul_rn(alpha, T);
                        const float4 q2 = lds128(ra + 32u), q3 = lds128(ra + 48u);
                        const float dj = q1.w - (dx * q2.x + dy * q2.y);
                        D = fmaf(dj, w, D);
                        C0 = fmaf(q2.z, w, C0); C1 = fmaf(q2.w, w, C1); C2 = fmaf(q3.x, w, C2);
                        N0 = fmaf(q3.y, w, N0); N1 = fmaf(q3.z, w, N1); N2 = fmaf(q3.w, w, N2);
                        T = test_T;
                        last = (uint32_t)(b * FWD_BULK + j + 1);
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
            __syncwarp();
            for (int c = 0; c < chunks; c++) {
                const uint32_t j = (uint32_t)(c * 32 + lane);
                const uint32_t mk = lds32(lm_warp + 4u * j);
                const unsigned hb = __ballot_sync(0xffffffffu, mk != 0u);
                if (mk != 0u)
                    hseg[hcnt + (uint32_t)__popc(hb & ((1u << lane) - 1u))] = make_uint2(lds32(rec_base + 64u * j + 8u), mk);
                hcnt += (uint32_t)__popc(hb);
            }
        }
        id_cur = id_b1;
        // closes the batch: its buffer and s_wm / s_lm may be overwritten afterwards; also the early-out vote
        if (__syncthreads_count(done) == EGS_TILE_THREADS) {
            if (b + 1 < nb) mbar_wait(bar_smem + 8u * (uint32_t)((b + 1) & 1), (uint32_t)(((b + 1) >> 1) & 1));   // drain
            break;
        }
    }
    if (lane == 0) im.hit_count[8 * tile + warp] = hcnt;
    if (inside) {
        T = fminf(0.999999f, T);
        im.final_T[pix] = T;
        im.final_D[pix] = D;
        im.n_contrib[pix] = last;
        out_color[pix] = fmaf(T, __ldg(bg), C0);
        out_color[HW + pix] = fmaf(T, __ldg(bg + 1), C1);
        out_color[2 * HW + pix] = fmaf(T, __ldg(bg + 2), C2);
        out_normal[pix] = N0; out_normal[HW + pix] = N1; out_normal[2 * HW + pix] = N2;
        out_depth[pix] = D / (1.f - T);
        out_opac[pix] = 1.f - T;
    }
}

cudaError_t launch_render_forward2(const egs_frame&, GeomView, ImgView, BinView, long long, float*, float*, float*,
                                   float*, bool, cudaStream_t);

// EGS_FWD_KERNEL: unset / "pair" = two pixels per lane on the packed FP32 pipe (egs_render_fwd2.cu, default);
// "ldg" = one pixel per lane (this file); "bulk" = one pixel per lane with TMA-staged batches (this file).
cudaError_t launch_render_forward(const egs_frame& f, GeomView g, ImgView im, BinView bn, long long cap,
                                  float* out_color, float* out_normal, float* out_depth, float* out_opac, bool save,
                                  cudaStream_t s) {
    const int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
    static int fwd_bulk = -1, fwd_pair = 1;
    if (fwd_bulk < 0) {
        const char* e = getenv("EGS_FWD_KERNEL");
        fwd_bulk = (e && e[0] == 'b') ? 1 : 0;
        fwd_pair = (e && (e[0] == 'b' || e[0] == 'l')) ? 0 : 1;
    }
    if (fwd_pair && (!save || egs_bwd_variant() == 3))
        return launch_render_forward2(f, g, im, bn, cap, out_color, out_normal, out_depth, out_opac, save, s);
    if (!save)
        k_render_forward<2><<<gx * gy, EGS_TILE_THREADS, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap, out_color,
                                                                 out_normal, out_depth, out_opac);
    else if (egs_bwd_variant() == 3 && fwd_bulk)
        k_render_forward_bulk<<<gx * gy, EGS_TILE_THREADS, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap,
                                                                   out_color, out_normal, out_depth, out_opac);
    else if (egs_bwd_variant() == 3)
        k_render_forward<1><<<gx * gy, EGS_TILE_THREADS, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap,
                                                                    out_color, out_normal, out_depth, out_opac);
    else
        k_render_forward<0><<<gx * gy, EGS_TILE_THREADS, 0, s>>>(f.width, f.height, gx, f.bg, g.rec, im, bn, cap,
                                                                     out_color, out_normal, out_depth, out_opac);
    return cudaGetLastError();
}
