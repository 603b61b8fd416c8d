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
       const int32_t* __restrict__ tile_mask, ImgView im, BinView bn, long long cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(rec + i));
    const float depth = __ldg(&rec[i].depth);
    int x0, y0, x1, y1;
    egs_tile_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
    const unsigned long long key = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned)i;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const int t = y * gx + x;
            if (tile_mask != nullptr && __ldg(tile_mask + t) == 0) continue;
            const long long pos = (long long)im.tile_offset[t] + atomicAdd(im.tile_cursor + t, 1u);
            if (pos < cap) bn.keys[pos] = key;
            else im.counters->overflow = 1;
        }
}

// ---- 4. per-tile sort ------------------------------------------------------------------------------------------
// Bitonic network in its all-ascending form (first step of every merge pairs i with i ^ (2k-1)), so the
// virtual padding beyond n never has to exist: a partner index >= n simply means "no exchange".
template <int CAP>
__global__ void __launch_bounds__(256) k_tile_sort(ImgView im, BinView bn, long long cap) {
    extern __shared__ unsigned long long skeys[];
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
    const bool in_smem = n <= CAP;
    unsigned long long* a = in_smem ? skeys : gk;
    if (in_smem) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = gk[i];
    }
    __syncthreads();
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
    k_emit<<<(P + 255) / 256, 256, 0, s>>>(P, gx, gy, radii, g.rec, tile_mask, im, bn, cap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    constexpr int CAP = 4096; // keys sorted in shared memory (32 KB); longer lists fall back to global memory
    k_tile_sort<CAP><<<gx * gy, 256, CAP * sizeof(unsigned long long), s>>>(im, bn, cap);
    return cudaGetLastError();
}
