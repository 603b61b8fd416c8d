// egs_binning.cu -- tile binning without a global sort.
//
// The reference builds one list of (tile << 32 | depth_bits, surfel) pairs and radix-sorts all of it
// (duplicateWithKeys + cub::DeviceRadixSort::SortPairs + identifyTileRanges + a host-side compaction,
// DGS/cuda_rasterizer/rasterizer_impl.cu:70-142,307-366).  Here the tile id is a bucket, not key bits:
//   1. k_surfel_forward already histogrammed instances per tile (tile_count);
//   2. k_tile_scan: exclusive scan of the <= 32k tile counts -> tile_offset (== the reference's `ranges`),
//      compacted ascending list of non-empty tiles (== `tile_indices`), num_rendered, tile_num;
//   3. k_emit: every visible surfel drops (depth_bits << 32 | id) into its tiles' segments (atomic cursor);
//   4. k_tile_sort: one CTA per tile sorts its segment by the 64-bit key in shared memory.
// Sorting by (depth bits, surfel id) gives exactly the order of the reference's stable sort: its ties
// (same tile, same depth bits) keep emission order, which is ascending surfel id.
#include "egs_common.cuh"

// ---- 2. scan -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += n;
    }
    return v;
}

__global__ void __launch_bounds__(1024) k_tile_scan(ImgView im, int tiles, long long cap) {
    __shared__ uint32_t warp_sum[2][32];
    __shared__ uint32_t carry[2];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) { carry[0] = 0; carry[1] = 0; }
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int i = base + t;
        const uint32_t c = i < tiles ? im.tile_count[i] : 0u;
        const uint32_t flag = c != 0u;
        uint32_t sc = warp_incl_scan(c), sf = warp_incl_scan(flag);
        if (lane == 31) { warp_sum[0][w] = sc; warp_sum[1][w] = sf; }
        __syncthreads();
        if (w == 0) {
            uint32_t a = warp_sum[0][lane], b = warp_sum[1][lane];
            a = warp_incl_scan(a); b = warp_incl_scan(b);
            warp_sum[0][lane] = a; warp_sum[1][lane] = b;
        }
        __syncthreads();
        const uint32_t pc = carry[0] + (w ? warp_sum[0][w - 1] : 0u);
        const uint32_t pf = carry[1] + (w ? warp_sum[1][w - 1] : 0u);
        if (i < tiles) {
            im.tile_offset[i] = pc + sc - c;   // exclusive
            im.tile_cursor[i] = 0u;
            if (flag) im.tile_list[pf + sf - 1] = i;
        }
        __syncthreads();
        if (t == 1023) { carry[0] = pc + sc; carry[1] = pf + sf; }
        __syncthreads();
    }
    // tail of the compacted list reads as -1, like the reference's tile_indices
    for (int i = carry[1] + t; i < tiles; i += 1024) im.tile_list[i] = -1;
    if (t == 0) {
        im.tile_offset[tiles] = carry[0];
        im.counters->num_rendered = (int32_t)carry[0];
        im.counters->tile_num = (int32_t)carry[1];
        if ((long long)carry[0] > cap && cap >= 0) im.counters->overflow = 1;
    }
}

// ---- 3. emit -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_emit(int P, int gx, int gy, const int32_t* __restrict__ radii, const SplatRecord* __restrict__ rec,
       const uint32_t* __restrict__ tiles_touched, const int32_t* __restrict__ tile_mask, ImgView im, BinView bn,
       long long cap) {
    // The warp's (surfel, tile) pairs are flattened and dealt to the lanes 32 at a time: every lane has one independent
    // cursor atomic in flight.  One thread walking its own rectangle was a serial chain of atomic round trips as long
    // as the warp's largest rectangle (ncu: 68 % of the kernel's stall samples sat on the returned cursor value).
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, y0 = 0, w = 0, area = 0;
    uint32_t depth_bits = 0u;
    if (i < P) {
        const int r = radii[i];
        // tiles_touched == 0: nothing to emit (all its tiles are masked out); in a sharded frame such a surfel may not
        // even have a record
        if (r > 0 && tiles_touched[i] != 0u) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(rec + i));
            depth_bits = __float_as_uint(__ldg(&rec[i].depth));
            int x1, y1;
            egs_tile_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
            w = x1 - x0;
            area = w > 0 && y1 > y0 ? w * (y1 - y0) : 0;
        }
    }
    // exclusive prefix of the areas over the warp
    int incl = area;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int off = incl - area;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int wpr = egs_mask_words_per_row(gx);
    const int warp_first = i - lane;
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        // source lane: the last one whose offset is <= t (offsets are non-decreasing; it always has area > 0 for t < total)
        int src = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int o = __shfl_sync(0xffffffffu, off, src + step);
            if (o <= t) src += step;
        }
        const int local = t - __shfl_sync(0xffffffffu, off, src);
        const int sw = __shfl_sync(0xffffffffu, w, src);
        const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t sdepth = __shfl_sync(0xffffffffu, depth_bits, src);
        if (t >= total) continue;
        const int ry = local / sw;
        const int x = sx0 + (local - ry * sw), y = sy0 + ry;
        if (tile_mask != nullptr && !egs_mask_bit(im.mask_bits, wpr, x, y)) continue;
        const int tile = y * gx + x;
        const long long pos = (long long)im.tile_offset[tile] + atomicAdd(im.tile_cursor + tile, 1u);
        if (pos < cap) bn.keys[pos] = ((unsigned long long)sdepth << 32) | (unsigned)(warp_first + src);
        else im.counters->overflow = 1;
    }
}

// ---- 4. per-tile sort ------------------------------------------------------------------------------------------
// Bitonic network in its all-ascending form: the first step of every merge of size k pairs element e with
// e ^ (k-1), the following steps pair e with e ^ j (j = k/4 ... 1); the smaller key always goes to the lower index.
// One CTA of 128 threads per tile.  Lists of up to 4096 keys are sorted in REGISTERS: thread t owns the E
// consecutive elements t*E .. t*E+E-1 (E = 2 ... 32 by list length, padded with +inf keys), so
//   * strides below E are compare-exchanges inside a thread (no communication, no barrier),
//   * strides E ... 16E exchange through warp shuffles (no barrier),
//   * only the few strides >= 32E of the last merges cross warps: the keys make one round trip through shared
//     memory per merge for those.
// A 512-key list thus needs 3 barriers-with-exchange instead of 45 (ncu before: 101 M warp-instructions, all
// generic loads/stores + barriers).  Longer lists (> 4096) fall back to the same network run in global memory.
#define SORT_THREADS 128
#define SORT_MAX_E 32

__device__ __forceinline__ void cex(unsigned long long& lo, unsigned long long& hi) {
    const unsigned long long a = lo, b = hi;
    const bool sw = a > b;
    lo = sw ? b : a;
    hi = sw ? a : b;
}

// one step (stride J) of the merge of size K on the register-resident keys
template <int E, int K, int J>
__device__ __forceinline__ void sort_step_regs(unsigned long long (&key)[E], int t) {
    constexpr bool first = (J == K / 2);
    if constexpr (J >= 32 * E) {
        // crosses warps: done in shared memory by the caller
    } else if constexpr (J >= E) {
        // partner thread: all lower thread bits flipped on the first step of a merge, one bit otherwise
        constexpr int tmask = first ? (K / E - 1) : (J / E);
        const bool keep_min = (t & (J / E)) == 0;
        if constexpr (first) {
            // the partner's element index is mirrored: my r meets its E-1-r.  Both halves of a mirrored pair are
            // fetched before either is overwritten.
#pragma unroll
            for (int r = 0; r < E / 2; r++) {
                const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, key[E - 1 - r], tmask);   // meets my r
                const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, key[r], tmask);           // meets my E-1-r
                const unsigned long long m1 = key[r], m2 = key[E - 1 - r];
                key[r] = keep_min ? (m1 < o1 ? m1 : o1) : (m1 > o1 ? m1 : o1);
                key[E - 1 - r] = keep_min ? (m2 < o2 ? m2 : o2) : (m2 > o2 ? m2 : o2);
            }
        } else {
#pragma unroll
            for (int r = 0; r < E; r++) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, key[r], tmask);
                const unsigned long long m = key[r];
                key[r] = keep_min ? (m < o ? m : o) : (m > o ? m : o);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < E; r++) {
            const int l = first ? (r ^ (K - 1)) : (r | J);
            if ((r & J) == 0 && l > r) cex(key[r], key[l]);
        }
    }
    if constexpr (J > 1) sort_step_regs<E, K, J / 2>(key, t);
}

template <int E, int K>
__device__ __forceinline__ void sort_merges_regs(unsigned long long (&key)[E], int t, int np2, unsigned long long* sm) {
    constexpr int NP2 = SORT_THREADS * E;
    if (K <= np2) {   // everything beyond np2 is padding: larger merges are no-ops
        if constexpr (K / 2 >= 32 * E) {
            // strides >= 32 E cross warps: one round trip through shared memory for all of them
            __syncthreads();
#pragma unroll
            for (int r = 0; r < E; r++) sm[t * E + r] = key[r];
            __syncthreads();
            for (int j = K >> 1; j >= 32 * E; j >>= 1) {
                const bool first = (j == (K >> 1));
                for (int p = t; p < NP2 / 2; p += SORT_THREADS) {
                    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                    const int l = first ? (i ^ (K - 1)) : (i | j);
                    const unsigned long long x = sm[i], y = sm[l];
                    if (x > y) { sm[i] = y; sm[l] = x; }
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < E; r++) key[r] = sm[t * E + r];
        }
        sort_step_regs<E, K, K / 2>(key, t);
        if constexpr (K < NP2) sort_merges_regs<E, K * 2>(key, t, np2, sm);
    }
}

template <int E>
__device__ __noinline__ void tile_sort_regs(const unsigned long long* __restrict__ gk, uint32_t* __restrict__ out,
                                            int n, unsigned long long* sm) {
    const int t = threadIdx.x;
    unsigned long long key[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int e = t * E + r;
        key[r] = e < n ? gk[e] : ~0ull;
    }
    int np2 = 2;
    while (np2 < n) np2 <<= 1;
    sort_merges_regs<E, 2>(key, t, np2, sm);
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int e = t * E + r;
        if (e < n) out[e] = (uint32_t)key[r];
    }
}

__global__ void __launch_bounds__(SORT_THREADS) k_tile_sort(ImgView im, BinView bn, long long cap) {
    __shared__ unsigned long long sm[SORT_THREADS * SORT_MAX_E];
    const int tile = blockIdx.x;
    const long long start = im.tile_offset[tile];
    long long end = im.tile_offset[tile + 1];
    if (end > cap) end = cap;
    const int n = (int)(end - start);
    if (n <= 0) return;
    unsigned long long* gk = bn.keys + start;
    uint32_t* out = bn.point_list + start;
    if (n == 1) {
        if (threadIdx.x == 0) out[0] = (uint32_t)gk[0];
        return;
    }
    if (n <= SORT_THREADS * 2) return tile_sort_regs<2>(gk, out, n, sm);
    if (n <= SORT_THREADS * 4) return tile_sort_regs<4>(gk, out, n, sm);
    if (n <= SORT_THREADS * 8) return tile_sort_regs<8>(gk, out, n, sm);
    if (n <= SORT_THREADS * 16) return tile_sort_regs<16>(gk, out, n, sm);
    if (n <= SORT_THREADS * 32) return tile_sort_regs<32>(gk, out, n, sm);
    // very long lists: the same network in global memory
    unsigned long long* a = gk;
    int np2 = 2;
    while (np2 < n) np2 <<= 1;
    const int half = np2 >> 1;
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool first = (j == (k >> 1));
            for (int p = threadIdx.x; p < half; p += blockDim.x) {
                // p-th compare-exchange of this step: i has bit j clear
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int l = first ? (i ^ (k - 1)) : (i | j);
                if (l < n) {
                    const unsigned long long x = a[i], y = a[l];
                    if (x > y) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = (uint32_t)a[i];
}

// ---- launchers ---------------------------------------------------------------------------------------------
cudaError_t launch_tile_scan(ImgView im, int tiles, long long cap, cudaStream_t s) {
    k_tile_scan<<<1, 1024, 0, s>>>(im, tiles, cap);
    return cudaGetLastError();
}

cudaError_t launch_emit_sort(int P, int gx, int gy, const int32_t* radii, GeomView g, const int32_t* tile_mask,
                             ImgView im, BinView bn, long long cap, cudaStream_t s) {
    if (P == 0) return cudaSuccess;
    k_emit<<<(P + 255) / 256, 256, 0, s>>>(P, gx, gy, radii, g.rec, g.tiles_touched, tile_mask, im, bn, cap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_tile_sort<<<gx * gy, SORT_THREADS, 0, s>>>(im, bn, cap);
    return cudaGetLastError();
}
