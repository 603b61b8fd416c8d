// egs_api.cu -- the extern "C" surface declared in include/eggsplat.h.  Argument checking, workspace carving and
// kernel sequencing only; no allocation, no device synchronisation, no exceptions.
#include "egs_common.cuh"
#include "../../include/eggmap.h"

cudaError_t launch_surfel_forward(const egs_frame&, const float*, const float*, const float*, const float*, const float*,
                                  const float*, const int32_t*, GeomView, ImgView, int32_t*, uint8_t*, int, int,
                                  cudaStream_t);
cudaError_t launch_surfel_backward(const egs_frame&, int, int, const float*, const float*, const float*, const float*,
                                   const float*, const int32_t*, GeomView, const float*, float*, float*, float*, float*,
                                   float*, float*, float*, float*, cudaStream_t);
cudaError_t launch_surfel_backward_adam(const egs_frame&, int, int, const float*, float*, const float*, const float*,
                                        const int32_t*, GeomView, const float*, float*, float*, float*, float*, float*,
                                        float*, double, double, double, int, float, float, cudaStream_t);
cudaError_t launch_mark_visible(int, const float*, const float*, const float*, uint8_t*, cudaStream_t);
cudaError_t launch_tile_scan(ImgView, int, long long, cudaStream_t);
cudaError_t launch_emit_sort(int, int, int, const int32_t*, GeomView, const int32_t*, ImgView, BinView, long long,
                             cudaStream_t);
cudaError_t launch_render_forward(const egs_frame&, GeomView, ImgView, BinView, long long, float*, float*, float*,
                                  float*, bool, cudaStream_t);
cudaError_t launch_render_backward(const egs_frame&, GeomView, ImgView, BinView, long long, const float*, const float*,
                                   const float*, const float*, float*, cudaStream_t);

cudaError_t launch_project_surfels(int, int, int, const float*, const float*, const uint8_t*, const float*, const float*,
                                   const float*, unsigned long long*, int32_t*, float*, cudaStream_t);
cudaError_t launch_fuse_surfels(int, int, int, const float*, const float*, const float*, const float*, const float*,
                                const float*, const uint8_t*, const int32_t*, float*, float*, float*, uint8_t*, uint8_t*,
                                float, float, float, cudaStream_t);

namespace {
inline int tiles_of(const egs_frame* f, int& gx, int& gy) {
    gx = (f->width + EGS_TILE - 1) / EGS_TILE;
    gy = (f->height + EGS_TILE - 1) / EGS_TILE;
    return gx * gy;
}
inline int check_frame(const egs_frame* f) {
    if (!f || f->num_surfels < 0 || f->width <= 0 || f->height <= 0) return EGS_E_BADARG;
    if (!f->bg || !f->viewmatrix || !f->projmatrix || !f->campos) return EGS_E_BADARG;
    if (f->sh_degree < 0 || f->sh_degree > 3) return EGS_E_UNSUPPORTED;
    return 0;
}
#define EGS_TRY(expr)                              \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

// Exchange step of a tile-sharded frame, over NVLink peer memory instead of a dense reduce-scatter.
// A surfel that touched one of this rank's tiles (tiles_touched != 0; at 8 GPUs ~1/5 of the visible ones) has a
// partial row in the rank's local screen-gradient block.  k_push_rows, one CTA per 256 consecutive surfels (all owned
// by one rank: chunk_rows is a multiple of 256): the touched rows are taken (and cleared, so the local block is all
// zeros again for the next step: no memset), stamped with their surfel id in the last padding word and compacted
// into shared memory; ONE counter bump per CTA reserves a run of slots in this sender's section of the OWNER's inbox,
// and the CTA streams the run there with fully coalesced 16-byte stores on the peer-mapped address (512 contiguous
// bytes per warp instruction).  k_push_counts publishes the counts; after a barrier the owner folds its inboxes
// into its (zeroed) block with local reductions (k_fold_inbox).
// What NVLink wants here was measured (2 x B200, NV18): bulk copy 714 GB/s; 64-byte rows scattered to random peer
// addresses 94 GB/s; red.global.add.v4.f32 of every quad onto the owner's block 71 GB/s (C3 push stage 0.25 ms);
// rows stored at their natural (55 % dense) positions 0.20 ms; warp-granular slot counters 0.12 - 0.29 ms (the
// same-address atomics serialise).  inbox layout on every rank: [world senders][chunk rows][16 floats], header int32 [world].
#define PUSH_CTA 256
__global__ void __launch_bounds__(PUSH_CTA)
k_push_rows(int P, int chunk, int rank, const uint32_t* __restrict__ tiles_touched, float* __restrict__ local_sg,
            uint32_t* __restrict__ sent, float* const* __restrict__ peer_inbox) {
    __shared__ __align__(128) float4 s_rows[PUSH_CTA * 4];
    __shared__ uint32_t s_warp[PUSH_CTA / 32];
    __shared__ uint32_t s_base;
    const int i = blockIdx.x * PUSH_CTA + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int owner = (blockIdx.x * PUSH_CTA) / chunk;      // uniform: chunk % PUSH_CTA == 0
    // The row is fetched together with the touched flag, not after it: a CTA lives for two dependent memory round trips
    // (flag + row | slot reservation + clear) instead of three.  The kernel is pure latency (ncu: 4 % of the issue slots
    // busy, 54 % of the stall samples on the row fetch behind the flag); reading the untouched rows as well costs
    // bandwidth it has to spare.
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
    float4* src = reinterpret_cast<float4*>(local_sg + (size_t)EGS_SCREEN_GRAD_STRIDE * (i < P ? i : 0));
    uint32_t touched = 0u;
    if (i < P) {
        touched = tiles_touched[i];
        v0 = src[0]; v1 = src[1]; v2 = src[2]; v3 = src[3];
    }
    const bool on = touched != 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_warp[warp] = (uint32_t)__popc(bal);
    __syncthreads();
    uint32_t before = 0u, total = 0u;
#pragma unroll
    for (int w = 0; w < PUSH_CTA / 32; w++) {
        const uint32_t c = s_warp[w];
        before += w < warp ? c : 0u;
        total += c;
    }
    if (total == 0u) return;
    if (threadIdx.x == 0) s_base = atomicAdd(sent + owner, total);
    if (on) {
        const uint32_t slot = before + (uint32_t)__popc(bal & ((1u << lane) - 1u));
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        src[0] = z; src[1] = z; src[2] = z; src[3] = z;
        v3.w = __int_as_float(i);
        s_rows[4 * slot] = v0; s_rows[4 * slot + 1] = v1; s_rows[4 * slot + 2] = v2; s_rows[4 * slot + 3] = v3;
    }
    // the compacted run leaves through the TMA engine: one bulk store of up to 16 KB per CTA on the peer address
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy writes -> async proxy
    __syncthreads();
#ifndef PUSH_NO_TMA
    if (threadIdx.x == 0) {
        float* dst = peer_inbox[owner] + ((size_t)rank * chunk + s_base) * EGS_SCREEN_GRAD_STRIDE;
        bulk_copy_s2g(dst, smem_addr(s_rows), 64u * total);
        bulk_commit_wait_read();   // the rows must stay in shared memory until the store has read them
    }
#else
    float4* dst = reinterpret_cast<float4*>(peer_inbox[owner] + ((size_t)rank * chunk + s_base) * EGS_SCREEN_GRAD_STRIDE);
    for (uint32_t k = threadIdx.x; k < 4u * total; k += PUSH_CTA) dst[k] = s_rows[k];
#endif
}

__global__ void k_push_counts(int world, int rank, uint32_t* __restrict__ sent, int32_t* const* __restrict__ peer_header) {
    const int r = threadIdx.x;
    if (r >= world) return;
    peer_header[r][rank] = (int32_t)sent[r];
    sent[r] = 0u;
}

// owner side: one thread per (slot, quad) of sender blockIdx.y's section
__global__ void __launch_bounds__(256)
k_fold_inbox(int chunk, int first, const float* __restrict__ inbox, const int32_t* __restrict__ header,
             float* __restrict__ block) {
    const int sender = blockIdx.y;
    const long long n4 = 4ll * header[sender];
    const float* sec = inbox + (size_t)sender * chunk * EGS_SCREEN_GRAD_STRIDE;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t >> 2;
        const int q = (int)(t & 3);
        const float* src = sec + row * EGS_SCREEN_GRAD_STRIDE;
        const int id = __float_as_int(__ldg(src + 15));
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
        float* dst = block + (size_t)EGS_SCREEN_GRAD_STRIDE * (id - first) + 4 * q;
        if (q < 3)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v.x) : "memory");
    }
}

__global__ void k_export_ranges(ImgView im, int tiles, uint32_t* ranges, int32_t* tile_indices) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tiles) return;
    const uint32_t a = im.tile_offset[t], b = im.tile_offset[t + 1];
    if (ranges) {   // the reference memsets ranges to 0 and only touches tiles that own instances
        ranges[2 * t] = a == b ? 0u : a;
        ranges[2 * t + 1] = a == b ? 0u : b;
    }
    if (tile_indices) tile_indices[t] = im.tile_list[t];
}
} // namespace

extern "C" {

EGS_API int egs_abi_version(void) { return EGS_ABI_VERSION; }

EGS_API const char* egs_error_string(int code) {
    if (code == 0) return "success";
    if (code == EGS_E_BADARG) return "eggsplat: bad argument";
    if (code == EGS_E_UNSUPPORTED) return "eggsplat: unsupported configuration";
    return cudaGetErrorString((cudaError_t)code);
}

EGS_API int egs_workspace_sizes(int32_t P, int32_t width, int32_t height, int64_t cap_instances, size_t* geom_bytes,
                        size_t* img_bytes, size_t* bin_bytes) {
    if (P < 0 || width <= 0 || height <= 0 || cap_instances < 0) return EGS_E_BADARG;
    const size_t tiles = (size_t)((width + EGS_TILE - 1) / EGS_TILE) * ((height + EGS_TILE - 1) / EGS_TILE);
    if (geom_bytes) *geom_bytes = carve_geom(nullptr, (size_t)P).bytes;
    if (img_bytes) *img_bytes = carve_img(nullptr, tiles, (size_t)width * height).bytes;
    if (bin_bytes) *bin_bytes = carve_bin(nullptr, (size_t)cap_instances).bytes;
    return 0;
}

EGS_API int egs_bin_bytes_forward_only(int64_t cap_instances, size_t* bin_bytes) {
    if (cap_instances < 0 || !bin_bytes) return EGS_E_BADARG;
    *bin_bytes = bin_bytes_forward_only((size_t)cap_instances);
    return 0;
}

EGS_API int egs_forward_plan(const egs_frame* f, const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* scales, const float* rotations, const int32_t* tile_mask,
                     void* geom, void* img, int32_t* radii, uint8_t* active_mask, egs_counters* counters_host,
                     void* stream) {
    return egs_forward_plan_sharded(f, means3D, shs, colors_precomp, opacities, scales, rotations, tile_mask, 0,
                                    f ? f->num_surfels : 0, geom, img, radii, active_mask, counters_host, stream);
}

EGS_API int egs_forward_plan_sharded(const egs_frame* f, const float* means3D, const float* shs,
                                     const float* colors_precomp, const float* opacities, const float* scales,
                                     const float* rotations, const int32_t* tile_mask, int32_t own_first,
                                     int32_t own_count, void* geom, void* img, int32_t* radii, uint8_t* active_mask,
                                     egs_counters* counters_host, void* stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    if (!img) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    int gx, gy;
    const int tiles = tiles_of(f, gx, gy);
    const int P = f->num_surfels;
    ImgView im = carve_img(img, (size_t)tiles, (size_t)f->width * f->height);
    // counters + ticket + tile_count are contiguous at the head of the image workspace
    EGS_TRY(cudaMemsetAsync(img, 0, (size_t)((char*)im.tile_offset - (char*)img), s));
    if (P > 0) {
        if (!means3D || !opacities || !scales || !rotations || !geom || !radii || !active_mask) return EGS_E_BADARG;
        if (!shs && !colors_precomp) return EGS_E_BADARG;
        if (!colors_precomp && f->sh_coeffs < (f->sh_degree + 1) * (f->sh_degree + 1)) return EGS_E_BADARG;
        GeomView g = carve_geom(geom, (size_t)P);
        if (own_first < 0 || own_count < 0 || own_first + own_count > P) return EGS_E_BADARG;
        EGS_TRY(launch_surfel_forward(*f, means3D, scales, rotations, opacities, shs, colors_precomp, tile_mask, g, im,
                                      radii, active_mask, own_first, own_count, s));
    }
    EGS_TRY(launch_tile_scan(im, tiles, -1, s));
    if (counters_host) EGS_TRY(cudaMemcpyAsync(counters_host, im.counters, sizeof(egs_counters), cudaMemcpyDeviceToHost, s));
    return 0;
}

EGS_API int egs_forward_render(const egs_frame* f, const int32_t* tile_mask, const int32_t* radii, void* geom, void* img,
                       void* bin, int64_t cap_instances, float* out_color, float* out_normal, float* out_depth,
                       float* out_opacity, egs_counters* counters_host, int32_t flags, void* stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    if (!img || !out_color || !out_normal || !out_depth || !out_opacity || cap_instances < 0) return EGS_E_BADARG;
    if (f->num_surfels > 0 && (!geom || !radii)) return EGS_E_BADARG;
    if (cap_instances > 0 && !bin) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    int gx, gy;
    const int tiles = tiles_of(f, gx, gy);
    const int P = f->num_surfels;
    ImgView im = carve_img(img, (size_t)tiles, (size_t)f->width * f->height);
    GeomView g = carve_geom(geom, (size_t)P);
    BinView bn = carve_bin(bin, (size_t)cap_instances);
    if (!(flags & EGS_FWD_REUSE_BINNING))
        EGS_TRY(launch_emit_sort(P, gx, gy, radii, g, tile_mask, im, bn, (long long)cap_instances, s));
    EGS_TRY(launch_render_forward(*f, g, im, bn, (long long)cap_instances, out_color, out_normal, out_depth,
                                  out_opacity, (flags & EGS_FWD_NO_SAVE) == 0, s));
    if (counters_host) EGS_TRY(cudaMemcpyAsync(counters_host, im.counters, sizeof(egs_counters), cudaMemcpyDeviceToHost, s));
    return 0;
}

EGS_API int egs_backward_render(const egs_frame* f, const void* geom, const void* img, const void* bin, int64_t cap_instances,
                        const float* dL_dcolor, const float* dL_dnormal, const float* dL_ddepth,
                        const float* dL_dopacity, float* screen_grads, int32_t flags, void* stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    const int P = f->num_surfels;
    if (P == 0) return 0;
    if (!geom || !img || !dL_dcolor || !dL_dnormal || !dL_ddepth || !dL_dopacity || !screen_grads) return EGS_E_BADARG;
    if (cap_instances > 0 && !bin) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    int gx, gy;
    const int tiles = tiles_of(f, gx, gy);
    ImgView im = carve_img(const_cast<void*>(img), (size_t)tiles, (size_t)f->width * f->height);
    GeomView g = carve_geom(const_cast<void*>(geom), (size_t)P);
    BinView bn = carve_bin(const_cast<void*>(bin), (size_t)cap_instances);
    if (!(flags & EGS_BWD_GRADS_PREZEROED))
        EGS_TRY(cudaMemsetAsync(screen_grads, 0, sizeof(float) * EGS_SCREEN_GRAD_STRIDE * (size_t)P, s));
    EGS_TRY(launch_render_backward(*f, g, im, bn, (long long)cap_instances, dL_dcolor, dL_dnormal, dL_ddepth,
                                   dL_dopacity, screen_grads, s));
    return 0;
}

EGS_API int egs_backward_surfels(const egs_frame* f, int32_t first, int32_t count, const float* means3D, const float* shs,
                         const float* colors_precomp, const float* scales, const float* rotations,
                         const int32_t* radii, const void* geom, const float* screen_grads, float* dL_dmeans3D,
                         float* dL_dopacity, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                         float* dL_dmeans2D, float* dL_dcolors, float* dL_dcov3D, void* stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    const int P = f->num_surfels;
    if (first < 0 || count < 0 || first + count > P) return EGS_E_BADARG;
    if (count == 0) return 0;
    if (!means3D || !scales || !rotations || !radii || !geom || !screen_grads) return EGS_E_BADARG;
    if (!dL_dmeans3D || !dL_dopacity || !dL_dscales || !dL_drotations) return EGS_E_BADARG;
    if (!colors_precomp && (!shs || !dL_dsh)) return EGS_E_BADARG;
    GeomView g = carve_geom(const_cast<void*>(geom), (size_t)P);
    EGS_TRY(launch_surfel_backward(*f, first, count, means3D, shs, colors_precomp, scales, rotations, radii, g,
                                   screen_grads, dL_dmeans3D, dL_dopacity, dL_dsh, dL_dscales, dL_drotations,
                                   dL_dmeans2D, dL_dcolors, dL_dcov3D, (cudaStream_t)stream));
    return 0;
}

EGS_API int egm_backward_surfels_adam(const egs_frame* f, int32_t first, int32_t count, const float* means3D, float* shs,
                                      const float* scales, const float* rotations, const int32_t* radii,
                                      const void* geom, const float* screen_grads, float* dL_dmeans3D,
                                      float* dL_dopacity, float* dL_dscales, float* dL_drotations,
                                      const egm_adam* hyper, float* m_shs, float* v_shs, void* stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    const int P = f->num_surfels;
    if (first < 0 || count < 0 || first + count > P || !hyper || hyper->step < 1) return EGS_E_BADARG;
    if (count == 0) return 0;
    if (!means3D || !shs || !scales || !rotations || !radii || !geom || !screen_grads || !m_shs || !v_shs) return EGS_E_BADARG;
    if (!dL_dmeans3D || !dL_dopacity || !dL_dscales || !dL_drotations) return EGS_E_BADARG;
    // the fused form exists for the layout the mapping loop uses: 16 SH coefficients, 16-byte aligned rows
    if (f->sh_coeffs != 16 || ((reinterpret_cast<uintptr_t>(shs) | reinterpret_cast<uintptr_t>(m_shs) |
                                reinterpret_cast<uintptr_t>(v_shs)) & 15) != 0)
        return EGS_E_UNSUPPORTED;
    GeomView g = carve_geom(const_cast<void*>(geom), (size_t)P);
    EGS_TRY(launch_surfel_backward_adam(*f, first, count, means3D, shs, scales, rotations, radii, g, screen_grads,
                                        dL_dmeans3D, dL_dopacity, dL_dscales, dL_drotations, m_shs, v_shs, hyper->beta1,
                                        hyper->beta2, hyper->eps, hyper->step, hyper->lr_f_dc, hyper->lr_f_rest,
                                        (cudaStream_t)stream));
    return 0;
}

EGS_API int egs_push_rows(int32_t P, int32_t chunk_rows, int32_t world, int32_t rank, const void* geom,
                          float* local_screen_grads, uint32_t* sent_counters, float* const* peer_inboxes,
                          int32_t* const* peer_headers, void* stream) {
    if (P < 0 || chunk_rows <= 0 || chunk_rows % PUSH_CTA != 0 || world <= 0 || world > 1024 || rank < 0 || rank >= world)
        return EGS_E_BADARG;
    if (!sent_counters || !peer_inboxes || !peer_headers) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (P > 0) {
        if (!geom || !local_screen_grads) return EGS_E_BADARG;
        GeomView g = carve_geom(const_cast<void*>(geom), (size_t)P);
        k_push_rows<<<(P + PUSH_CTA - 1) / PUSH_CTA, PUSH_CTA, 0, s>>>(P, chunk_rows, rank, g.tiles_touched,
                                                                       local_screen_grads, sent_counters, peer_inboxes);
        EGS_TRY(cudaGetLastError());
    }
    k_push_counts<<<1, 1024, 0, s>>>(world, rank, sent_counters, peer_headers);
    EGS_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egs_fold_inbox(int32_t chunk_rows, int32_t world, int32_t first, const float* inbox, const int32_t* header,
                           float* block, void* stream) {
    if (chunk_rows <= 0 || world <= 0 || first < 0 || !inbox || !header || !block) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    EGS_TRY(cudaMemsetAsync(block, 0, sizeof(float) * EGS_SCREEN_GRAD_STRIDE * (size_t)chunk_rows, s));
    const int bx = (int)((4ll * chunk_rows + 255) / 256 < 1184 ? (4ll * chunk_rows + 255) / 256 : 1184);   // 148 SMs x 8
    k_fold_inbox<<<dim3((unsigned)bx, (unsigned)world), 256, 0, s>>>(chunk_rows, first, inbox, header, block);
    EGS_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
    if (P < 0) return EGS_E_BADARG;
    if (P == 0) return 0;
    if (!means3D || !viewmatrix || !projmatrix || !present) return EGS_E_BADARG;
    EGS_TRY(launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream));
    return 0;
}

EGS_API int egs_project_surfels(int32_t P, int32_t height, int32_t width, const float* points, const float* rotations,
                                const uint8_t* stable_mask, const float* intrinsic, const float* viewmatrix,
                                const float* projmatrix, void* scratch, int32_t* index_map, float* depth_buffer,
                                void* stream) {
    if (P < 0 || height <= 0 || width <= 0) return EGS_E_BADARG;
    if (!scratch || !index_map || !depth_buffer || !intrinsic || !viewmatrix || !projmatrix) return EGS_E_BADARG;
    if (P > 0 && (!points || !rotations || !stable_mask)) return EGS_E_BADARG;
    EGS_TRY(launch_project_surfels(P, height, width, points, rotations, stable_mask, intrinsic, viewmatrix, projmatrix,
                                   (unsigned long long*)scratch, index_map, depth_buffer, (cudaStream_t)stream));
    return 0;
}

EGS_API int egs_fuse_surfels(int32_t P, int32_t height, int32_t width, const float* intrinsic, const float* viewmatrix,
                             const float* projmatrix, const float* frame_vmap, const float* frame_nmap,
                             const float* frame_dmap, const uint8_t* frame_mask, const int32_t* frame_imap,
                             float* points, float* rotations, float* sigma2, uint8_t* inview_mask,
                             uint8_t* surface_mask, float fusion_dist_thres, float alpha_p, float alpha_n,
                             void* stream) {
    if (P < 0 || height <= 0 || width <= 0) return EGS_E_BADARG;
    if (P == 0) return 0;
    if (!intrinsic || !viewmatrix || !projmatrix || !frame_vmap || !frame_nmap || !frame_dmap || !frame_mask ||
        !frame_imap || !points || !rotations || !sigma2 || !inview_mask || !surface_mask)
        return EGS_E_BADARG;
    EGS_TRY(launch_fuse_surfels(P, height, width, intrinsic, viewmatrix, projmatrix, frame_vmap, frame_nmap, frame_dmap,
                                frame_mask, frame_imap, points, rotations, sigma2, inview_mask, surface_mask,
                                fusion_dist_thres, alpha_p, alpha_n, (cudaStream_t)stream));
    return 0;
}

EGS_API int egs_debug_export(const egs_frame* f, const void* geom, const void* img, const void* bin, int64_t cap_instances,
                     uint32_t* point_list, uint32_t* ranges, int32_t* tile_indices, uint32_t* tiles_touched,
                     uint32_t* n_contrib, float* final_T, float* final_D, float* records, float* cov3D,
                     uint8_t* clamped, void* stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    int gx, gy;
    const int tiles = tiles_of(f, gx, gy);
    const size_t P = (size_t)f->num_surfels, N = (size_t)f->width * f->height;
    ImgView im = carve_img(const_cast<void*>(img), (size_t)tiles, N);
    const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
    if (ranges || tile_indices) {
        k_export_ranges<<<(tiles + 255) / 256, 256, 0, s>>>(im, tiles, ranges, tile_indices);
        EGS_TRY(cudaGetLastError());
    }
    if (n_contrib) EGS_TRY(cudaMemcpyAsync(n_contrib, im.n_contrib, 4 * N, d2d, s));
    if (final_T) EGS_TRY(cudaMemcpyAsync(final_T, im.final_T, 4 * N, d2d, s));
    if (final_D) EGS_TRY(cudaMemcpyAsync(final_D, im.final_D, 4 * N, d2d, s));
    if (P > 0 && geom) {
        GeomView g = carve_geom(const_cast<void*>(geom), P);
        if (tiles_touched) EGS_TRY(cudaMemcpyAsync(tiles_touched, g.tiles_touched, 4 * P, d2d, s));
        if (records) EGS_TRY(cudaMemcpyAsync(records, g.rec, sizeof(SplatRecord) * P, d2d, s));
        if (cov3D) EGS_TRY(cudaMemcpyAsync(cov3D, g.cov3D, 24 * P, d2d, s));
        if (clamped) EGS_TRY(cudaMemcpyAsync(clamped, g.clamped, P, d2d, s));
    }
    if (point_list && bin && cap_instances > 0) {
        BinView bn = carve_bin(const_cast<void*>(bin), (size_t)cap_instances);
        EGS_TRY(cudaMemcpyAsync(point_list, bn.point_list, 4 * (size_t)cap_instances, d2d, s));
    }
    return 0;
}

} // extern "C"
