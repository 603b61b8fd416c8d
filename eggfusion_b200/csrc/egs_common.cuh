// egs_common.cuh -- shared definitions for the sm_100a surfel rasterizer kernels.
//
// HBM layout (all caller-owned, see include/eggsplat.h):
//   geom workspace, per surfel:   SplatRecord rec[P] (64 B, 16-B aligned quads) | cov3D[P][6] f32 |
//                                 tiles_touched[P] u32 | clamped[P] u8 | cand[P] i32 (sharded projection's candidate list)
//   img  workspace:               egs_counters (+ticket) | tile_count[T] | tile_offset[T+1] | tile_cursor[T] |
//                                 tile_list[T] (compacted non-empty tiles) | hit_count[8T] | mask_bits | final_T[N] | final_D[N] |
//                                 n_contrib[N]
//   bin  workspace, per instance: keys[cap] u64 (depth bits << 32 | surfel id, bucketed by tile) | point_list[cap] u32 |
//                                 hits[8*cap] {surfel id, pixel mask}: per (tile, 8x4 warp block) the compacted, depth-
//                                 ordered list of splats that blended into the block (block b of a tile whose list is
//                                 [start, start+n) owns entries 8*start + b*n ...); the tile-wide backward variants
//                                 use the same bytes as lane_masks[cap][8] u32 instead
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eggsplat.h"

#define EGS_TILE 16          // tile edge in pixels (reference: config.h:15-17)
#define EGS_TILE_THREADS 256 // one thread per pixel of a tile

#if defined(__CUDACC__)
#define EGS_HD __host__ __device__ __forceinline__
#else
#define EGS_HD inline
#endif

// ---- explicitly rounded fp32 primitives -------------------------------------------------------------------
// The index-critical chain (cull tests, radii, tile rectangles, depth keys) is written with these so that no
// compiler is free to contract or re-associate it: the device build uses the *_rn intrinsics, the host build
// (tests/hostemu) plain IEEE operations compiled with -ffp-contract=off.  The grouping mirrors the FFMA
// pattern nvcc 12.9 emits for the reference on sm_100a (see oracle/splat_oracle.c header).
#if defined(__CUDA_ARCH__)
EGS_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
EGS_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
EGS_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
EGS_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
EGS_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
EGS_HD float f_rcp(float a) { return __frcp_rn(a); }
EGS_HD float f_sqrt(float a) { return __fsqrt_rn(a); }
#else
#include <math.h>
EGS_HD float f_mul(float a, float b) { return a * b; }
EGS_HD float f_add(float a, float b) { return a + b; }
EGS_HD float f_sub(float a, float b) { return a - b; }
EGS_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
EGS_HD float f_div(float a, float b) { return a / b; }
EGS_HD float f_rcp(float a) { return 1.0f / a; }
EGS_HD float f_sqrt(float a) { return sqrtf(a); }
#endif

// a*b + c*d + e*f : second product rounded on its own, first and third fused
EGS_HD float f_dot3(float a, float b, float c, float d, float e, float f) {
    return f_fma(e, f, f_fma(a, b, f_mul(c, d)));
}

// ---- per-frame constants staged once per CTA --------------------------------------------------------------
struct FrameConst {
    float view[16];
    float proj[16];
    float campos[3];
    float bg[3];
    float tanfovx, tanfovy, fx, fy, cx, cy, mod;
    int W, H, gx, gy, D, M;
};

// 64-byte packed splat record: everything the two compositing kernels read per (tile, surfel) instance.
// q0 is enough for the per-warp bounding-box reject, q0+q1 for alpha, q2+q3 only for contributing pixels.
struct __align__(16) SplatRecord {
    float x, y;          // q0: pixel-space centre (means2D)
    uint32_t ext;        //     conservative half extents of the alpha >= 1/255 ellipse, 2 x u16 in 1/8 px (x | y << 16)
    float opacity;
    float cxx, cxy, cyy; // q1: conic
    float depth;         //     view-space z (also the sort key)
    float ja, jb;        // q2: plane-depth slope  d(depth)/d(pixel offset)  (Jinv0*u0z+Jinv2*u1z, Jinv1*u0z+Jinv3*u1z)
    float r, g;
    float b;             // q3
    float nx, ny, nz;    //     view-space normal (un-normalised, as the reference blends it)
};
static_assert(sizeof(SplatRecord) == 64, "record must be 64 bytes");

// ---- workspace carving --------------------------------------------------------------------------------------
EGS_HD size_t egs_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct GeomView {
    SplatRecord* rec;
    float* cov3D;
    uint32_t* tiles_touched;
    uint8_t* clamped;
    int32_t* cand;          // [P] sharded projection: ids of the surfels that may reach the rank's tiles (egs_preprocess.cu)
    size_t bytes;
};
struct ImgView {
    egs_counters* counters; // followed by ticket words
    uint32_t* ticket;
    uint32_t* tile_count;
    uint32_t* tile_offset; // [T + 1]
    uint32_t* tile_cursor;
    int32_t* tile_list;    // [T]
    uint32_t* hit_count;   // [8T] entries of each (tile, warp block) hit list
    uint32_t* mask_bits;   // [gy][ceil(gx/32)] the frame's tile mask, one bit per tile (rows start at a word; k_pack_mask)
    float* final_T;
    float* final_D;
    uint32_t* n_contrib;
    size_t bytes;
};
struct BinView {
    unsigned long long* keys;
    uint32_t* point_list;
    uint32_t* lane_masks; // [cap][8]: bit l of word w = pixel l of warp block w blended this instance in the forward
    uint2* hits;          // [8*cap] (aliases lane_masks): per-block hit lists {surfel id, mask of the block's pixels}
    size_t bytes;
};

EGS_HD GeomView carve_geom(void* base, size_t P) {
    GeomView v;
    size_t o = 0;
    char* b = (char*)base;
    v.rec = (SplatRecord*)(b + o);          o = egs_align_up(o + sizeof(SplatRecord) * P, 256);
    v.cov3D = (float*)(b + o);              o = egs_align_up(o + sizeof(float) * 6 * P, 256);
    v.tiles_touched = (uint32_t*)(b + o);   o = egs_align_up(o + sizeof(uint32_t) * P, 256);
    v.clamped = (uint8_t*)(b + o);          o = egs_align_up(o + P, 256);
    v.cand = (int32_t*)(b + o);             o = egs_align_up(o + sizeof(int32_t) * (P + 2048), 256);   // 32 sub-lists
    v.bytes = o + 256;
    return v;
}
EGS_HD ImgView carve_img(void* base, size_t tiles, size_t npix) {
    ImgView v;
    size_t o = 0;
    char* b = (char*)base;
    v.counters = (egs_counters*)(b + o);
    v.ticket = (uint32_t*)(b + 64);         o = 256;
    v.tile_count = (uint32_t*)(b + o);      o = egs_align_up(o + 4 * tiles, 256);
    v.tile_offset = (uint32_t*)(b + o);     o = egs_align_up(o + 4 * (tiles + 1), 256);
    v.tile_cursor = (uint32_t*)(b + o);     o = egs_align_up(o + 4 * tiles, 256);
    v.tile_list = (int32_t*)(b + o);        o = egs_align_up(o + 4 * tiles, 256);
    v.hit_count = (uint32_t*)(b + o);       o = egs_align_up(o + 32 * tiles, 256);
    v.mask_bits = (uint32_t*)(b + o);       o = egs_align_up(o + 4 * tiles, 256);   // gy * ceil(gx / 32) <= tiles words
    v.final_T = (float*)(b + o);            o = egs_align_up(o + 4 * npix, 256);
    v.final_D = (float*)(b + o);            o = egs_align_up(o + 4 * npix, 256);
    v.n_contrib = (uint32_t*)(b + o);       o = egs_align_up(o + 4 * npix, 256);
    v.bytes = o + 256;
    return v;
}
EGS_HD BinView carve_bin(void* base, size_t cap) {
    BinView v;
    size_t o = 0;
    char* b = (char*)base;
    v.keys = (unsigned long long*)(b + o);  o = egs_align_up(o + 8 * cap, 256);
    v.point_list = (uint32_t*)(b + o);      o = egs_align_up(o + 4 * cap, 256);
    v.lane_masks = (uint32_t*)(b + o);
    v.hits = (uint2*)(b + o);               o = egs_align_up(o + 64 * cap, 256);
    v.bytes = o + 256;
    return v;
}
// size of the prefix a forward-only render touches (keys + point list)
EGS_HD size_t bin_bytes_forward_only(size_t cap) {
    return egs_align_up(egs_align_up(8 * cap, 256) + 4 * cap, 256) + 256;
}

// getRect of the reference (auxiliary.h:47-57): float arithmetic, truncation toward zero, clamped to the grid.
EGS_HD void egs_tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1, int& y1) {
    const float r = (float)radius;
    int a;
    a = (int)f_mul(f_sub(px, r), 0.0625f);                                   x0 = a < 0 ? 0 : (a > gx ? gx : a);
    a = (int)f_mul(f_sub(py, r), 0.0625f);                                   y0 = a < 0 ? 0 : (a > gy ? gy : a);
    a = (int)f_mul(f_sub(f_add(f_add(px, r), 16.0f), 1.0f), 0.0625f);        x1 = a < 0 ? 0 : (a > gx ? gx : a);
    a = (int)f_mul(f_sub(f_add(f_add(py, r), 16.0f), 1.0f), 0.0625f);        y1 = a < 0 ? 0 : (a > gy ? gy : a);
}

// ---- bit-packed tile mask (sharded frames): 1 KB at 1080p, so every lookup is an L1 hit, where the caller's int32 mask
// (32 KB) kept missing behind the streaming loads of the per-surfel kernels
EGS_HD int egs_mask_words_per_row(int gx) { return (gx + 31) >> 5; }
#if defined(__CUDACC__)
// bits of row y's tiles [x0, x1) that lie in word xw (x0 < x1, xw in [x0 >> 5, (x1 - 1) >> 5])
__device__ __forceinline__ uint32_t egs_mask_row_word(const uint32_t* __restrict__ bits, int wpr, int y, int xw, int x0,
                                                      int x1) {
    const int lo = max(x0 - (xw << 5), 0), hi = min(x1 - (xw << 5), 32);
    const uint32_t range = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
    return __ldg(bits + y * wpr + xw) & range;
}
__device__ __forceinline__ bool egs_mask_bit(const uint32_t* __restrict__ bits, int wpr, int x, int y) {
    return (__ldg(bits + y * wpr + (x >> 5)) >> (x & 31)) & 1u;
}
#endif

// Which reverse-walk kernel runs (EGS_BWD_KERNEL, see egs_render_bwd.cu): 3 = warp (default, consumes per-block hit
// lists), 0/1/2 = tile-wide variants (consume lane_masks).  The forward writes the format the backward will read.
int egs_bwd_variant();

#if defined(__CUDACC__)
// 16-byte shared-memory load from a 32-bit shared-window address.  Indexing __shared__ arrays through generic
// pointers inside the hit loops made ptxas rebuild the window base (S2UR SR_CgaCtaId + UMOV + ULEA) per iteration.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a)); // opaque: keep it in a register instead of rematerialising the window base
    return a;
}

// ---- mbarrier + bulk copy (the TMA engine's non-tensor path; SASS UBLKCP / SYNCS) ---------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void bulk_copy_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Stage the per-frame constants into shared memory (one pass, first 64 threads).
__device__ __forceinline__ void load_frame_const(FrameConst& fc, const egs_frame& f) {
    const int t = threadIdx.x;
    if (t < 16) fc.view[t] = __ldg(f.viewmatrix + t);
    else if (t < 32) fc.proj[t - 16] = __ldg(f.projmatrix + t - 16);
    else if (t < 35) fc.campos[t - 32] = __ldg(f.campos + t - 32);
    else if (t < 38) fc.bg[t - 35] = __ldg(f.bg + t - 35);
    else if (t == 38) {
        fc.tanfovx = f.tanfovx; fc.tanfovy = f.tanfovy;
        fc.fy = f.height / (2.0f * f.tanfovy);   // rasterizer_impl.cu:244-245
        fc.fx = f.width / (2.0f * f.tanfovx);
        fc.cx = f.cx; fc.cy = f.cy; fc.mod = f.scale_modifier;
        fc.W = f.width; fc.H = f.height;
        fc.gx = (f.width + EGS_TILE - 1) / EGS_TILE; fc.gy = (f.height + EGS_TILE - 1) / EGS_TILE;
        fc.D = f.sh_degree; fc.M = f.sh_coeffs;
    }
}
#endif
