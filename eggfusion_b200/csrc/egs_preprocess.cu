// egs_preprocess.cu -- per-surfel kernels: forward projection (+ per-tile instance counting), per-surfel backward,
// and the coarse visibility test.  One thread per surfel, 128-thread CTAs; the CTA's SH block and the per-frame
// constants are staged in shared memory.
//
// Replaces preprocessCUDA<3> (DGS/cuda_rasterizer/forward.cu:158-301), computeCov2DCUDA + preprocessCUDA<3> (bwd)
// (DGS/cuda_rasterizer/backward.cu:144-416) and checkFrustum (DGS/cuda_rasterizer/rasterizer_impl.cu:54-66).
#include "egs_surfel_math.cuh"
#include "egm_math.cuh"
#include <stdlib.h>

#define SURF_THREADS 128
#define SH_PITCH 49   // floats per staged SH row (48 + 1: conflict-free 4-byte accesses, one row per thread)

// Coalesced copy of this CTA's `rows` SH rows (48 floats each, contiguous in global memory) into shared memory.
__device__ __forceinline__ void stage_sh_rows(float* s_sh, const float* __restrict__ src, int rows) {
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int idx = threadIdx.x; idx < rows * 12; idx += SURF_THREADS) {
        const float4 v = __ldg(src4 + idx);
        float* d = s_sh + (idx / 12) * SH_PITCH + 4 * (idx % 12);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}
__device__ __forceinline__ void unstage_sh_rows(const float* s_sh, float* __restrict__ dst, int rows) {
    float4* dst4 = reinterpret_cast<float4*>(dst);
    for (int idx = threadIdx.x; idx < rows * 12; idx += SURF_THREADS) {
        const float* d = s_sh + (idx / 12) * SH_PITCH + 4 * (idx % 12);
        dst4[idx] = make_float4(d[0], d[1], d[2], d[3]);
    }
}

// Fused optimiser step of the mapping iteration (egm_backward_surfels_adam, include/eggmap.h): the CTA's dL/dSH block is
// consumed where it was produced -- Adam on the SH rows straight out of shared memory -- instead of being written to
// HBM and read back by a separate pass (192 B/surfel each way, plus one launch).  m == nullptr: plain backward.
struct ShAdam {
    float* p;      // SH parameters [P][16][3] (the array the kernel's `shs` input points at), updated in place
    float* m;      // exp_avg
    float* v;      // exp_avg_sq
    EgmAdamConst c;
    float nss_dc, nss_rest;   // -(lr / (1 - beta1^t)) of row 0 (f_dc) and of rows 1..15 (f_rest)
};
// The parameter quad is re-read from global memory (the staged copy was overwritten in place by the gradient): the
// CTA fetched it a few microseconds ago, so it is an L2 hit.
__device__ __forceinline__ void adam_sh_rows(const float* s_sh, const ShAdam& a, size_t row0, int rows) {
    float4* p4 = reinterpret_cast<float4*>(a.p + 48 * row0);
    float4* m4 = reinterpret_cast<float4*>(a.m + 48 * row0);
    float4* v4 = reinterpret_cast<float4*>(a.v + 48 * row0);
#pragma unroll 4
    for (int idx = threadIdx.x; idx < rows * 12; idx += SURF_THREADS) {
        const int q = idx % 12;
        const float* d = s_sh + (idx / 12) * SH_PITCH + 4 * q;
        float4 p = p4[idx], m = m4[idx], v = v4[idx];
        // columns 0..2 of a row are f_dc (SH coefficient 0, three channels), the rest f_rest
        p.x = egm_adam_update(p.x, d[0], m.x, v.x, a.c, q == 0 ? a.nss_dc : a.nss_rest);
        p.y = egm_adam_update(p.y, d[1], m.y, v.y, a.c, q == 0 ? a.nss_dc : a.nss_rest);
        p.z = egm_adam_update(p.z, d[2], m.z, v.z, a.c, q == 0 ? a.nss_dc : a.nss_rest);
        p.w = egm_adam_update(p.w, d[3], m.w, v.w, a.c, a.nss_rest);
        p4[idx] = p; m4[idx] = m; v4[idx] = v;
    }
}

#define SH_BULK_PITCH 52   // floats per row for the bulk-copied layout: 208-B rows keep 16-B alignment and make the
                           // per-thread LDS.128 of a quarter-warp conflict-free (52 mod 32 = 20 -> 8 distinct 4-bank groups)

// Sharded frames only: can this surfel reach one of the rank's tiles at all?  surfel_bound_rect (egs_surfel_math.cuh)
// gives a conservative tile rectangle from the projected centre and an upper bound of the splat radius, so that the
// ~1700 instructions of the exact projection are spent only on the ~1/world of the surfels that can matter here.
__device__ __forceinline__ bool surfel_misses_mask(const FrameConst& fc, const float* mean, const float* scale,
                                                   const float* rot, const uint32_t* __restrict__ mask_bits) {
    int x0, y0, x1, y1;
    if (!surfel_bound_rect(fc, mean, scale, rot, x0, y0, x1, y1)) return false;   // undecided: leave it to the exact path
    const int w = x1 - x0, h = y1 - y0;
    if (w <= 0 || h <= 0) return true;                    // the bound rectangle is off the grid
    if (w * h > 48) return false;                         // a huge splat: not worth walking the mask
    // no early exit: the loads of one surfel are independent and pipeline (this pass is latency-, not issue-bound)
    const int wpr = egs_mask_words_per_row(fc.gx);
    uint32_t any = 0;
    for (int y = y0; y < y1; y++)
        for (int xw = x0 >> 5; xw <= (x1 - 1) >> 5; xw++) any |= egs_mask_row_word(mask_bits, wpr, y, xw, x0, x1);
    return any == 0u;
}

// one bit per tile of the caller's int32 tile mask (rows padded to whole words)
__global__ void k_pack_mask(const int32_t* __restrict__ tile_mask, int gx, int gy, uint32_t* __restrict__ bits) {
    const int wpr = egs_mask_words_per_row(gx);
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= wpr * gy) return;
    const int y = w / wpr, xb = (w % wpr) << 5;
    uint32_t word = 0;
    for (int b = 0; b < 32 && xb + b < gx; b++) word |= (__ldg(tile_mask + y * gx + xb + b) != 0 ? 1u : 0u) << b;
    bits[w] = word;
}

// Pass 1 of a sharded projection: surfels that cannot reach the rank's tiles (and are not owned) get zero radii /
// active / tiles_touched here and are never looked at again; the others are compacted into `cand` (warp-aggregated
// append, order irrelevant) so that the exact projection runs on dense warps.  Surfel order carries no spatial
// coherence, so skipping inside the projection kernel itself saves nothing (a warp always contains a surfel that
// needs the full path: measured, 0.110 -> 0.158 ms at 2 ranks).
// The list is kept in CAND_REGIONS sub-lists (warp w of the surfel array appends to region w % CAND_REGIONS, which can
// never hold more than its share of the surfels): one append counter per region instead of one for everything -- the
// warps of this kernel spent half their time queued behind ~31 k atomics on a single word (ncu), and a CTA-wide
// aggregation traded that for three barriers per group.
#define CAND_REGIONS 32
__host__ __device__ __forceinline__ int cand_region_cap(int P) {
    const int warps = (P + 31) / 32;
    return (warps + CAND_REGIONS - 1) / CAND_REGIONS * 32;
}

__global__ void __launch_bounds__(256)
k_surfel_candidates(const egs_frame f, const float* __restrict__ means, const float* __restrict__ scales,
                    const float* __restrict__ rots, const uint32_t* __restrict__ mask_bits, int own_first, int own_count, int32_t* __restrict__ radii,
                    uint8_t* __restrict__ active, uint32_t* __restrict__ tiles_touched, int32_t* __restrict__ cand,
                    int32_t* __restrict__ cand_count) {
    __shared__ FrameConst fc;
    load_frame_const(fc, f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int cap = cand_region_cap(f.num_surfels);
    // persistent CTAs, grid-stride over 256-surfel groups: the frame constants are fetched once per CTA
    for (int i0 = blockIdx.x * blockDim.x; i0 < f.num_surfels; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        if (i < f.num_surfels) {
            keep = (i >= own_first && i - own_first < own_count) ||
                   !surfel_misses_mask(fc, means + (size_t)3 * i, scales + (size_t)3 * i, rots + (size_t)4 * i, mask_bits);
            if (!keep) {
                radii[i] = 0;
                active[i] = 0;
                tiles_touched[i] = 0u;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (bal == 0u) continue;
        const int region = (i >> 5) & (CAND_REGIONS - 1);
        const int leader = __ffs(bal) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(cand_count + region, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) cand[(size_t)region * cap + base + __popc(bal & ((1u << lane) - 1u))] = i;
    }
}

// SH_SMEM = 1: the CTA's SH block (M == 16: 128 x 192 B, contiguous) is staged through shared memory with fully
// coalesced 16-byte loads and read on demand, instead of living in 48 registers per thread.
// SH_SMEM = 2: every thread issues ONE 192-byte bulk copy (cp.async.bulk -> mbarrier, the TMA engine) of its SH row at
// kernel start and waits for it only after the projection / culling math, so the fetch overlaps ~1000 instructions
// of arithmetic and costs one instruction instead of 12 LDG + 48 STS.
template <int SH_SMEM>
__global__ void __launch_bounds__(SURF_THREADS)
k_surfel_forward(const egs_frame f, const float* __restrict__ means, const float* __restrict__ scales,
                 const float* __restrict__ rots, const float* __restrict__ opac, const float* __restrict__ shs,
                 const float* __restrict__ colors, const int32_t* __restrict__ tile_mask, GeomView g, ImgView im,
                 int32_t* __restrict__ radii, uint8_t* __restrict__ active, int own_first, int own_count,
                 const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_count) {
    // own_first / own_count: the surfel range this rank owns in a tile-sharded frame.  A visible surfel whose
    // rectangle has no tile in the rank's mask and which lies outside the range is needed by nobody here: its colour
    // (SH evaluation, 192 B of coefficients) and its record / cov3D / clamp state are skipped.  SH_SMEM == 3 is the
    // staging for that case: the row's bulk copy is issued only once the thread knows it needs it.
    __shared__ FrameConst fc;
    __shared__ __align__(128) float s_sh[SH_SMEM >= 2 ? SURF_THREADS * SH_BULK_PITCH : (SH_SMEM ? SURF_THREADS * SH_PITCH : 1)];
    __shared__ __align__(8) unsigned long long s_bar;
    // cand != nullptr (sharded frames): thread j works on surfel cand[j], j < *cand_count -- the surfels a cheap
    // footprint bound could not rule out (k_surfel_candidates); everything else was zeroed there
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (cand) {
        // CTA (region, q) of the sub-list layout k_surfel_candidates wrote
        const int cap = cand_region_cap(f.num_surfels);
        const int per_region = (cap + SURF_THREADS - 1) / SURF_THREADS;
        const int region = blockIdx.x / per_region, j = (blockIdx.x % per_region) * SURF_THREADS + threadIdx.x;
        const int n_work = cand_count[region];
        if (j - (int)threadIdx.x >= n_work) return;             // whole CTA beyond the sub-list (uniform)
        i = j < n_work ? cand[(size_t)region * cap + j] : f.num_surfels;
    }
    load_frame_const(fc, f);
    const uint32_t bar = smem_addr(&s_bar);
    if (SH_SMEM == 1) {
        const int row0 = blockIdx.x * SURF_THREADS;
        stage_sh_rows(s_sh, shs + (size_t)48 * row0, min(SURF_THREADS, f.num_surfels - row0));
    }
    if ((SH_SMEM == 2 || SH_SMEM == 3) && threadIdx.x == 0) {
        mbar_init(bar, SH_SMEM == 2 ? 1u : (uint32_t)SURF_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (SH_SMEM == 2) {
        const int row0 = blockIdx.x * SURF_THREADS;
        const int rows = min(SURF_THREADS, f.num_surfels - row0);
        if (threadIdx.x == 0) mbar_expect_tx(bar, 192u * (uint32_t)rows);
        if ((int)threadIdx.x < rows)
            bulk_copy_g2s(smem_addr(s_sh) + 4u * SH_BULK_PITCH * threadIdx.x, shs + (size_t)48 * (row0 + threadIdx.x), 192u, bar);
    }
    const bool valid = i < f.num_surfels;
    // candidate list: 9 in 10 of its surfels end up needing their SH row, so it is fetched up front and arrives behind the
    // projection math (deferring it exposed the whole fetch latency: 20 % of the kernel's stall samples)
    const bool early = SH_SMEM == 3 && cand != nullptr;
    if (early) {
        if (valid) {
            mbar_expect_tx(bar, 192u);
            bulk_copy_g2s(smem_addr(s_sh) + 4u * SH_BULK_PITCH * threadIdx.x, shs + (size_t)48 * i, 192u, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    bool visible = false, need = false;
    SurfelFwd o;
    uint32_t cnt = 0;
    const bool use_sh = colors == nullptr;
    const float* mean = means + (size_t)3 * (valid ? i : 0);
    if (valid) {
        surfel_forward(fc, mean, scales + (size_t)3 * i, rots + (size_t)4 * i, __ldg(opac + i), o);
        radii[i] = o.radius;
        active[i] = (uint8_t)o.active;
        if (o.radius > 0) {
            visible = true;
            // per-tile instance counts: the histogram that replaces the reference's per-surfel scan
            if (tile_mask == nullptr) {
                for (int y = o.y0; y < o.y1; y++)
                    for (int x = o.x0; x < o.x1; x++) atomicAdd(im.tile_count + y * fc.gx + x, 1u);
                cnt = (uint32_t)((o.y1 - o.y0) * (o.x1 - o.x0));
            } else if (o.x1 > o.x0) {
                const int wpr = egs_mask_words_per_row(fc.gx);
                for (int y = o.y0; y < o.y1; y++)
                    for (int xw = o.x0 >> 5; xw <= (o.x1 - 1) >> 5; xw++) {
                        uint32_t m = egs_mask_row_word(im.mask_bits, wpr, y, xw, o.x0, o.x1);
                        cnt += __popc(m);
                        while (m) {
                            atomicAdd(im.tile_count + y * fc.gx + (xw << 5) + __ffs(m) - 1, 1u);
                            m &= m - 1u;
                        }
                    }
            }
            need = cnt > 0u || (i >= own_first && i - own_first < own_count);
        }
        g.tiles_touched[i] = cnt;
    }
    if (SH_SMEM == 3 && !early) {
        // every thread arrives exactly once, before anybody waits: the ones that need their SH row add its bytes
        if (need) {
            mbar_expect_tx(bar, 192u);
            bulk_copy_g2s(smem_addr(s_sh) + 4u * SH_BULK_PITCH * threadIdx.x, shs + (size_t)48 * i, 192u, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    if (need) {
        if (SH_SMEM >= 2) {
            mbar_wait(bar, 0u);
            surfel_color(fc, mean, s_sh + threadIdx.x * SH_BULK_PITCH, true, o);
        } else if (SH_SMEM == 1) {
            surfel_color(fc, mean, s_sh + threadIdx.x * SH_PITCH, true, o);
        } else if (use_sh) {
            // generic layout: this surfel's 3*(D+1)^2 floats into registers (16-byte vectors when rows allow it)
            float shreg[48];
            const float* src = shs + (size_t)3 * fc.M * i;
            const int nfl = 3 * (fc.D + 1) * (fc.D + 1);
            if (((3 * fc.M) & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) {
#pragma unroll
                for (int q = 0; q < 12; q++)
                    if (4 * q < nfl) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
                        shreg[4 * q] = v.x; shreg[4 * q + 1] = v.y; shreg[4 * q + 2] = v.z; shreg[4 * q + 3] = v.w;
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 48; k++)
                    if (k < nfl) shreg[k] = __ldg(src + k);
            }
            surfel_color(fc, mean, shreg, true, o);
        } else {
            surfel_color(fc, mean, colors + (size_t)3 * i, false, o);
        }
        float4* dst = reinterpret_cast<float4*>(g.rec + i);
        const float4* src = reinterpret_cast<const float4*>(&o.rec);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        float2* c2 = reinterpret_cast<float2*>(g.cov3D + (size_t)6 * i);
        c2[0] = make_float2(o.cov3D[0], o.cov3D[1]);
        c2[1] = make_float2(o.cov3D[2], o.cov3D[3]);
        c2[2] = make_float2(o.cov3D[4], o.cov3D[5]);
        g.clamped[i] = (uint8_t)o.clamped;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, visible);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&im.counters->num_visible, __popc(ballot));
    if (SH_SMEM >= 2) mbar_wait(bar, 0u);   // the CTA must not retire while bulk copies into its shared memory are in flight
}

// ------------------------------------------------------------------------------------------------ backward
// SH_SMEM = 2: the SH row of every VISIBLE surfel arrives by one 192-byte bulk copy issued at kernel start (culled
// rows are not fetched at all) and dL/dSH leaves by one bulk store per row; no CTA barrier around either.
template <int SH_SMEM>
__global__ void __launch_bounds__(SURF_THREADS)
k_surfel_backward(const egs_frame f, int first, int count, const float* __restrict__ means,
                  const float* __restrict__ shs, const float* __restrict__ colors, const float* __restrict__ scales,
                  const float* __restrict__ rots, const int32_t* __restrict__ radii, GeomView g,
                  const float* __restrict__ sg, float* __restrict__ d_means, float* __restrict__ d_opacity,
                  float* __restrict__ d_sh, float* __restrict__ d_scales, float* __restrict__ d_rots,
                  float* __restrict__ d_means2D, float* __restrict__ d_colors, float* __restrict__ d_cov3D,
                  const ShAdam adam) {
    __shared__ FrameConst fc;
    __shared__ __align__(128) float s_sh[SH_SMEM == 2 ? SURF_THREADS * SH_BULK_PITCH : (SH_SMEM ? SURF_THREADS * SH_PITCH : 1)];
    __shared__ __align__(8) unsigned long long s_bar;
    constexpr int PITCH = SH_SMEM == 2 ? SH_BULK_PITCH : SH_PITCH;
    load_frame_const(fc, f);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int row0 = first + blockIdx.x * SURF_THREADS;
    const int rows = min(SURF_THREADS, first + count - row0);
    const uint32_t bar = smem_addr(&s_bar);
    if (SH_SMEM == 1) stage_sh_rows(s_sh, shs + (size_t)48 * row0, rows);
    if (SH_SMEM == 2 && threadIdx.x == 0) {
        mbar_init(bar, SURF_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool valid = k < count;
    const int i = first + k;
    const int M = fc.M;
    const bool use_sh = colors == nullptr;
    const uint32_t row_smem = smem_addr(s_sh) + 4u * PITCH * threadIdx.x;
    if (SH_SMEM == 2) {
        // every thread arrives once; the threads of visible surfels add their row's bytes and start the copy
        if (valid && radii[i] > 0) {
            mbar_expect_tx(bar, 192u);
            bulk_copy_g2s(row_smem, shs + (size_t)48 * i, 192u, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    if (valid) {
        float g16[16];
        SurfelBwd o;
        const bool vis = radii[i] > 0;
        float* my_sh = use_sh ? d_sh + (size_t)3 * M * i : nullptr;
        float* row = s_sh + threadIdx.x * PITCH;
        if (vis) {
            const float4* grow = reinterpret_cast<const float4*>(sg + (size_t)EGS_SCREEN_GRAD_STRIDE * i);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 v = __ldg(grow + q);
                g16[4 * q] = v.x; g16[4 * q + 1] = v.y; g16[4 * q + 2] = v.z; g16[4 * q + 3] = v.w;
            }
            const float* mean = means + (size_t)3 * i;
            surfel_backward_geom(fc, mean, scales + (size_t)3 * i, rots + (size_t)4 * i, g.cov3D + (size_t)6 * i, g16, o);
            if (use_sh) {
                const float dir[3] = {mean[0] - fc.campos[0], mean[1] - fc.campos[1], mean[2] - fc.campos[2]};
                const float gcol[3] = {g16[6], g16[7], g16[8]};
                float add[3];
                if (SH_SMEM) {
                    if (SH_SMEM == 2) mbar_wait(bar, 0u);
                    // in place: sh_backward finishes reading the row before its first store
                    sh_backward(fc.D, row, dir, (uint32_t)g.clamped[i], gcol,
                                [row](int kk, int ch, float v) { row[3 * kk + ch] = v; }, add);
                    const int used = 3 * (fc.D + 1) * (fc.D + 1);
                    for (int kk = used; kk < 48; kk++) row[kk] = 0.f; // coefficients above the active degree
                } else {
                    const int nfl = 3 * (fc.D + 1) * (fc.D + 1);
                    const bool vec = ((3 * M) & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0 &&
                                     (reinterpret_cast<uintptr_t>(d_sh) & 15) == 0;
                    float shreg[48], dsh[48];
                    const float* src = shs + (size_t)3 * M * i;
                    if (vec) {
#pragma unroll
                        for (int q = 0; q < 12; q++)
                            if (4 * q < nfl) {
                                const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
                                shreg[4 * q] = v.x; shreg[4 * q + 1] = v.y; shreg[4 * q + 2] = v.z; shreg[4 * q + 3] = v.w;
                            }
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 48; kk++)
                            if (kk < nfl) shreg[kk] = __ldg(src + kk);
                    }
#pragma unroll
                    for (int kk = 0; kk < 48; kk++) dsh[kk] = 0.f;
                    sh_backward(fc.D, shreg, dir, (uint32_t)g.clamped[i], gcol,
                                [&dsh](int kk, int ch, float v) { dsh[3 * kk + ch] = v; }, add);
                    if (vec) {
#pragma unroll
                        for (int q = 0; q < 12; q++)
                            if (4 * q < 3 * M)
                                reinterpret_cast<float4*>(my_sh)[q] = make_float4(dsh[4 * q], dsh[4 * q + 1], dsh[4 * q + 2], dsh[4 * q + 3]);
                        for (int kk = 48; kk < 3 * M; kk++) my_sh[kk] = 0.f;
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 48; kk++)
                            if (kk < 3 * M) my_sh[kk] = dsh[kk];
                        for (int kk = 48; kk < 3 * M; kk++) my_sh[kk] = 0.f;
                    }
                }
                o.d_mean[0] += add[0]; o.d_mean[1] += add[1]; o.d_mean[2] += add[2];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; q++) g16[q] = 0.f;
#pragma unroll
            for (int q = 0; q < 3; q++) { o.d_mean[q] = 0.f; o.d_scale[q] = 0.f; }
#pragma unroll
            for (int q = 0; q < 4; q++) o.d_rot[q] = 0.f;
#pragma unroll
            for (int q = 0; q < 6; q++) o.d_cov3D[q] = 0.f;
            if (use_sh) {
                if (SH_SMEM) {
                    for (int kk = 0; kk < 48; kk++) row[kk] = 0.f;
                } else if (((3 * M) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_sh) & 15) == 0) {
                    for (int q = 0; 4 * q < 3 * M; q++) reinterpret_cast<float4*>(my_sh)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    for (int kk = 0; kk < 3 * M; kk++) my_sh[kk] = 0.f;
                }
            }
        }
        d_means[3 * (size_t)i] = o.d_mean[0]; d_means[3 * (size_t)i + 1] = o.d_mean[1]; d_means[3 * (size_t)i + 2] = o.d_mean[2];
        d_scales[3 * (size_t)i] = o.d_scale[0]; d_scales[3 * (size_t)i + 1] = o.d_scale[1]; d_scales[3 * (size_t)i + 2] = o.d_scale[2];
        reinterpret_cast<float4*>(d_rots)[i] = make_float4(o.d_rot[0], o.d_rot[1], o.d_rot[2], o.d_rot[3]);
        d_opacity[i] = g16[5];
        if (d_means2D) { d_means2D[3 * (size_t)i] = g16[0]; d_means2D[3 * (size_t)i + 1] = g16[1]; d_means2D[3 * (size_t)i + 2] = 0.f; }
        if (d_colors) { d_colors[3 * (size_t)i] = g16[6]; d_colors[3 * (size_t)i + 1] = g16[7]; d_colors[3 * (size_t)i + 2] = g16[8]; }
        if (d_cov3D) {
#pragma unroll
            for (int q = 0; q < 6; q++) d_cov3D[6 * (size_t)i + q] = o.d_cov3D[q];
        }
    }
    if (SH_SMEM == 1) {
        __syncthreads();
        if (adam.m != nullptr) adam_sh_rows(s_sh, adam, (size_t)row0, rows);
        else unstage_sh_rows(s_sh, d_sh + (size_t)48 * row0, rows); // coalesced 16-byte stores of the CTA's dL/dSH block
    }
    if (SH_SMEM == 2) {
        mbar_wait(bar, 0u);   // no copy into this CTA's shared memory may be in flight when it retires
        if (valid) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy row writes -> async proxy
            bulk_copy_s2g(d_sh + (size_t)48 * i, row_smem, 192u);
        }
        bulk_commit_wait_read();   // the row must stay in shared memory until the store has read it
    }
}

// ------------------------------------------------------------------------------------------------ markVisible
__global__ void __launch_bounds__(256)
k_mark_visible(int P, const float* __restrict__ means, const float* __restrict__ view, const float* __restrict__ proj,
               uint8_t* __restrict__ present) {
    __shared__ float sv[16], sp[16];
    if (threadIdx.x < 16) sv[threadIdx.x] = view[threadIdx.x];
    else if (threadIdx.x < 32) sp[threadIdx.x - 16] = proj[threadIdx.x - 16];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float x = means[3 * (size_t)i], y = means[3 * (size_t)i + 1], z = means[3 * (size_t)i + 2];
    const float hx = xf_affine(sp, 0, x, y, z), hy = xf_affine(sp, 1, x, y, z), hw = xf_affine(sp, 3, x, y, z);
    const float pw = f_rcp(f_add(hw, 0.0000001f));
    const float ndx = f_mul(hx, pw), ndy = f_mul(hy, pw);
    const float vz = xf_affine(sv, 2, x, y, z);
    // auxiliary.h:168: the +-1.3 literals are doubles
    present[i] = !(vz <= 0.2f || (double)ndx < -1.3 || (double)ndx > 1.3 || (double)ndy < -1.3 || (double)ndy > 1.3);
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_surfel_forward(const egs_frame& f, const float* means, const float* scales, const float* rots,
                                  const float* opac, const float* shs, const float* colors, const int32_t* tile_mask,
                                  GeomView g, ImgView im, int32_t* radii, uint8_t* active, int own_first, int own_count,
                                  cudaStream_t s) {
    const int P = f.num_surfels;
    if (P == 0) return cudaSuccess;
    // staged-SH fast path: SH colours with exactly 16 coefficients per surfel and 16-byte aligned rows
    const bool sh_smem = colors == nullptr && f.sh_coeffs == 16 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0;
    // EGS_SH_STAGE: "ldg" = LDG + STS staging, "bulk" = TMA bulk copies; default = bulk here (measured at C3:
    // 96 -> 87 us, the fetch overlaps the projection math), ldg in the backward (bulk measured equal: 114 vs 115 us)
    static int sh_bulk = -1;
    if (sh_bulk < 0) {
        const char* e = getenv("EGS_SH_STAGE");
        sh_bulk = (e && e[0] == 'l') ? 0 : 1;
    }
    if (tile_mask != nullptr) {
        // every kernel of the frame looks the mask up through its bit-packed copy (egs_common.cuh)
        int gx = (f.width + EGS_TILE - 1) / EGS_TILE, gy = (f.height + EGS_TILE - 1) / EGS_TILE;
        const int words = egs_mask_words_per_row(gx) * gy;
        k_pack_mask<<<(words + 127) / 128, 128, 0, s>>>(tile_mask, gx, gy, im.mask_bits);
    }
    const bool sharded = tile_mask != nullptr && own_count < P;   // most SH rows will not be needed: fetch on demand
    if (sh_smem && sharded) {
        // two passes when the rank owns a third of the surfels or less (>= 3 ranks; measured at 2 ranks: the candidate
        // pass costs more than the projection of the ~25 % of the surfels it rules out): candidate compaction
        // (k_surfel_candidates), then the exact projection on the candidates only, on dense warps
        static int two_pass = -1;
        if (two_pass < 0) {
            const char* e = getenv("EGS_SHARD_TWO_PASS");   // "0": never, "1": always, default: own_count <= P / 3
            two_pass = e ? (e[0] == '1' ? 1 : 0) : 2;
        }
        if (two_pass == 1 || (two_pass == 2 && (long long)own_count * 3 <= (long long)P + 768)) {
            int32_t* cand_count = reinterpret_cast<int32_t*>(im.ticket);   // CAND_REGIONS words, zeroed by the plan's head memset
            const int groups = (P + 255) / 256;
            k_surfel_candidates<<<groups < 148 * 8 ? groups : 148 * 8, 256, 0, s>>>(
                f, means, scales, rots, im.mask_bits, own_first, own_count, radii, active, g.tiles_touched, g.cand, cand_count);
            const int per_region = (cand_region_cap(P) + SURF_THREADS - 1) / SURF_THREADS;
            k_surfel_forward<3><<<CAND_REGIONS * per_region, 128, 0, s>>>(f, means, scales, rots, opac, shs, colors,
                                                                          tile_mask, g, im, radii, active, own_first,
                                                                          own_count, g.cand, cand_count);
        } else
            k_surfel_forward<3><<<(P + 127) / 128, 128, 0, s>>>(f, means, scales, rots, opac, shs, colors, tile_mask, g,
                                                                im, radii, active, own_first, own_count, nullptr, nullptr);
    } else if (sh_smem && sh_bulk)
        k_surfel_forward<2><<<(P + 127) / 128, 128, 0, s>>>(f, means, scales, rots, opac, shs, colors, tile_mask, g, im,
                                                            radii, active, own_first, own_count, nullptr, nullptr);
    else if (sh_smem)
        k_surfel_forward<1><<<(P + 127) / 128, 128, 0, s>>>(f, means, scales, rots, opac, shs, colors, tile_mask, g, im,
                                                            radii, active, own_first, own_count, nullptr, nullptr);
    else
        k_surfel_forward<0><<<(P + 127) / 128, 128, 0, s>>>(f, means, scales, rots, opac, shs, colors, tile_mask, g,
                                                                im, radii, active, own_first, own_count, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_surfel_backward(const egs_frame& f, int first, int count, const float* means, const float* shs,
                                   const float* colors, const float* scales, const float* rots, const int32_t* radii,
                                   GeomView g, const float* sg, float* d_means, float* d_opacity, float* d_sh,
                                   float* d_scales, float* d_rots, float* d_means2D, float* d_colors, float* d_cov3D,
                                   cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const bool sh_smem = colors == nullptr && f.sh_coeffs == 16 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(d_sh) & 15) == 0;
    static int sh_bulk = -1;
    if (sh_bulk < 0) {
        const char* e = getenv("EGS_SH_STAGE");
        sh_bulk = (e && e[0] == 'b') ? 1 : 0;
    }
    ShAdam none;
    none.p = none.m = none.v = nullptr;
    if (sh_smem && sh_bulk)
        k_surfel_backward<2><<<(count + 127) / 128, 128, 0, s>>>(f, first, count, means, shs, colors, scales, rots,
                                                                 radii, g, sg, d_means, d_opacity, d_sh, d_scales,
                                                                 d_rots, d_means2D, d_colors, d_cov3D, none);
    else if (sh_smem)
        k_surfel_backward<1><<<(count + 127) / 128, 128, 0, s>>>(f, first, count, means, shs, colors, scales, rots,
                                                                    radii, g, sg, d_means, d_opacity, d_sh, d_scales,
                                                                    d_rots, d_means2D, d_colors, d_cov3D, none);
    else
        k_surfel_backward<0><<<(count + 127) / 128, 128, 0, s>>>(f, first, count, means, shs, colors, scales, rots,
                                                                     radii, g, sg, d_means, d_opacity, d_sh, d_scales,
                                                                     d_rots, d_means2D, d_colors, d_cov3D, none);
    return cudaGetLastError();
}

// per-surfel backward + Adam on the SH block in one kernel; shs is read AND updated (16 coefficients, 16-byte aligned)
cudaError_t launch_surfel_backward_adam(const egs_frame& f, int first, int count, const float* means, float* shs,
                                        const float* scales, const float* rots, const int32_t* radii, GeomView g,
                                        const float* sg, float* d_means, float* d_opacity, float* d_scales,
                                        float* d_rots, float* m_sh, float* v_sh, double beta1, double beta2, double eps,
                                        int step, float lr_dc, float lr_rest, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    ShAdam a;
    a.p = shs; a.m = m_sh; a.v = v_sh;
    // torch computes the bias corrections and the step size in double on the host (optim/adam.py); same as egm_adam_step
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    a.c.beta1 = (float)beta1; a.c.beta2 = (float)beta2;
    a.c.one_m_beta1 = (float)(1.0 - beta1); a.c.one_m_beta2 = (float)(1.0 - beta2);
    a.c.eps = (float)eps; a.c.bc2_sqrt = (float)sqrt(bc2);
    a.nss_dc = (float)(-((double)lr_dc / bc1));
    a.nss_rest = (float)(-((double)lr_rest / bc1));
    k_surfel_backward<1><<<(count + 127) / 128, 128, 0, s>>>(f, first, count, means, shs, nullptr, scales, rots, radii, g,
                                                             sg, d_means, d_opacity, nullptr, d_scales, d_rots, nullptr,
                                                             nullptr, nullptr, a);
    return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float* means, const float* view, const float* proj, uint8_t* present,
                                cudaStream_t s) {
    if (P == 0) return cudaSuccess;
    k_mark_visible<<<(P + 255) / 256, 256, 0, s>>>(P, means, view, proj, present);
    return cudaGetLastError();
}
