/*
 * eggsplat.h -- C ABI of libeggsplat.so: the B200 (sm_100a) surfel rasterizer hot path of EGG-Fusion.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes and a CUDA stream
 * (passed as void* so the header needs no CUDA include); nothing here allocates, frees, synchronises the
 * device or throws.  Return value: 0 on success, otherwise a cudaError_t value (or a negative EGS_E_* code
 * for argument errors); the text is available from egs_error_string().
 *
 * What each function replaces in the reference (/root/reference/submodules/diff-gaussian-surfels, "DGS"):
 *
 *   egs_workspace_sizes      GeometryState/ImageState/BinningState::fromChunk + required<T>()
 *                            (DGS/cuda_rasterizer/rasterizer_impl.cu:159-208, rasterizer_impl.h:74-83) and the
 *                            resize callbacks of DGS/rasterize_points.cu:27-33,82-87
 *   egs_forward_plan         first half of CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:212-311):
 *                            FORWARD::preprocess + the instance count
 *   egs_forward_render       second half (rasterizer_impl.cu:313-400): duplicateWithKeys, SortPairs,
 *                            identifyTileRanges, tile compaction, FORWARD::render
 *   egs_backward_render      BACKWARD::render (rasterizer_impl.cu:461-493, backward.cu:419-676)
 *   egs_backward_surfels     BACKWARD::preprocess (rasterizer_impl.cu:499-522, backward.cu:144-416)
 *   egs_mark_visible         Rasterizer::markVisible (rasterizer_impl.cu:145-157)
 *   egs_project_surfels      projectSurfelsToFrame / ProjectSurfelsToFrameCUDA (DGS/fuse_surfels.cu:475-573)
 *   egs_fuse_surfels         preprocessSurfel / PreprocessSurfelsFunctionCUDA (DGS/fuse_surfels.cu:214-471)
 *   egs_debug_export         (none: test hook that dumps the internal index artefacts for parity checks)
 *
 * Together egs_forward_plan + egs_forward_render are what `_C.rasterize_gaussians` binds
 * (DGS/ext.cpp:15-21, DGS/rasterize_points.cu:35-133); egs_backward_render + egs_backward_surfels are
 * `_C.rasterize_gaussians_backward` (rasterize_points.cu:135-229).  INTEGRATION.md shows the binding.
 */
#ifndef EGGSPLAT_H_
#define EGGSPLAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGS_ABI_VERSION 1

#if defined(__GNUC__)
#define EGS_API __attribute__((visibility("default")))
#else
#define EGS_API
#endif

#define EGS_E_BADARG (-1)      /* null / inconsistent argument */
#define EGS_E_UNSUPPORTED (-2) /* e.g. sh_degree > 3, or a path the reference itself cannot execute */

/* Floats per row of the screen-space gradient block G[P][16] that egs_backward_render accumulates and
 * egs_backward_surfels consumes (the block that is reduce-scattered between GPUs):
 *   0,1 mean2D.xy | 2,3,4 conic (xx, xy, yy) | 5 opacity | 6,7,8 colour | 9,10,11 normal | 12 depth | 13-15 zero */
#define EGS_SCREEN_GRAD_STRIDE 16

/* flags of egs_forward_render */
#define EGS_FWD_REUSE_BINNING 1 /* skip emission + sort: composite again from the lists already in `bin` */
#define EGS_FWD_NO_SAVE 2       /* forward-only render (the reference under torch.no_grad(), e.g. the model maps of
                                   tracking / fusion, mapper.py:227,497): nothing is kept for a backward -- no per-block hit
                                   lists, no saved per-pixel state -- and `bin` needs only egs_bin_bytes_forward_only() */
/* flags of egs_backward_render */
#define EGS_BWD_GRADS_PREZEROED 1 /* caller already zeroed (or pre-loaded) screen_grads: accumulate on top */

/* Scalars of GaussianRasterizationSettings (DGS/diff_gaussian_rasterization/__init__.py:166-180) plus
 * DEVICE pointers to its four small tensors, which are read on the device (no host copy, no sync). */
typedef struct egs_frame {
    int32_t num_surfels;  /* P */
    int32_t width, height;
    int32_t sh_degree;    /* active degree D, 0..3 */
    int32_t sh_coeffs;    /* M = coefficients stored per surfel in `shs` (ignored when colors_precomp != NULL) */
    float tanfovx, tanfovy;
    float cx, cy;
    float scale_modifier;
    const float* bg;         /* device [3]  */
    const float* viewmatrix; /* device [16] world_view_transform, row-major of W2C^T */
    const float* projmatrix; /* device [16] full_proj_transform */
    const float* campos;     /* device [3]  */
} egs_frame;

/* Device-resident counters written by egs_forward_plan / egs_forward_render (first bytes of the image workspace). */
typedef struct egs_counters {
    int32_t num_rendered; /* I: number of (surfel, tile) instances = what the reference returns as `rendered` */
    int32_t tile_num;     /* number of tiles with a non-empty list */
    int32_t overflow;     /* 1 if I exceeded the capacity of the binning workspace (lists were truncated) */
    int32_t num_visible;  /* number of surfels with radii > 0 */
} egs_counters;

EGS_API int egs_abi_version(void);
EGS_API const char* egs_error_string(int code);

/* Bytes needed for the three caller-owned workspaces.  `cap_instances` is the capacity (in instances) the
 * binning workspace is sized for; pass the exact count read back after egs_forward_plan, or an upper bound. */
EGS_API int egs_workspace_sizes(int32_t num_surfels, int32_t width, int32_t height, int64_t cap_instances,
                        size_t* geom_bytes, size_t* img_bytes, size_t* bin_bytes);

/* Bytes of the binning workspace when egs_forward_render is called with EGS_FWD_NO_SAVE (12 instead of 76 bytes per
 * instance: sort keys + point list only). */
EGS_API int egs_bin_bytes_forward_only(int64_t cap_instances, size_t* bin_bytes);

/*
 * Per-surfel projection.  Writes radii[P] (0 = culled), active_mask[P] (1 = inside the frustum), the packed
 * 64-byte splat records and per-surfel state into `geom`, per-tile instance counts, their exclusive scan, the
 * compacted list of non-empty tiles and the counters into `img`.  `tile_mask` ([tiles_y][tiles_x] int32, may be
 * NULL = all ones) restricts instance emission exactly like the reference's tile_mask.
 * `counters_host` (pinned host memory, may be NULL) receives an async copy of the counters on `stream`.
 */
EGS_API int egs_forward_plan(const egs_frame* frame, const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* scales, const float* rotations, const int32_t* tile_mask,
                     void* geom, void* img, int32_t* radii, uint8_t* active_mask, egs_counters* counters_host,
                     void* stream);

/*
 * egs_forward_plan for one rank of a tile-sharded frame (the seam is the reference's tile_mask,
 * forward.cu:292-300, rasterizer_impl.cu:103-111).  The rank renders the tiles of `tile_mask` and owns the surfels
 * [own_first, own_first + own_count) for the per-surfel backward.  The colour (SH evaluation), the splat record and the
 * per-surfel backward state of a visible surfel are produced only if it touches one of the rank's tiles or lies in the
 * owned range -- nobody on this rank reads them otherwise -- and a surfel that a cheap conservative bound of its
 * footprint places outside all of the rank's tiles (and outside the owned range) is not projected at all: its radii /
 * active_mask entries are 0 on THIS rank (the union over the ranks is the single-GPU result).
 * egs_forward_plan == own range [0, P): every entry exact.
 */
EGS_API int egs_forward_plan_sharded(const egs_frame* frame, const float* means3D, const float* shs,
                                     const float* colors_precomp, const float* opacities, const float* scales,
                                     const float* rotations, const int32_t* tile_mask, int32_t own_first,
                                     int32_t own_count, void* geom, void* img, int32_t* radii, uint8_t* active_mask,
                                     egs_counters* counters_host, void* stream);

/*
 * Exchange step of a tile-sharded frame over peer memory (one process per GPU, NVLink / NVSwitch), sender side: moves
 * the rows of `local_screen_grads` [P][16] that this rank's reverse walk touched into this sender's section of their
 * owners' inboxes (compacted per 256-surfel group in shared memory and sent with one bulk store -- cp.async.bulk,
 * the TMA engine -- on the peer-mapped address; the surfel id travels in the row's last padding word), clears them
 * locally and publishes the per-owner row counts.
 * Owner r owns surfels [r * chunk_rows, (r+1) * chunk_rows); chunk_rows must be a multiple of 256.
 * peer_inboxes / peer_headers: DEVICE arrays of `world` pointers to rank r's inbox (float [world][chunk_rows][16]) and
 * header (int32 [world]) as THIS process maps them.  sent_counters: device uint32 [world], zero before the first call
 * (the call leaves it zero).
 * The caller orders it against the owners' readers: a barrier after this call on every rank, then egs_fold_inbox.
 * Replaces a dense NCCL reduce-scatter of the whole block: only touched rows cross the links.
 */
EGS_API int egs_push_rows(int32_t num_surfels, int32_t chunk_rows, int32_t world, int32_t rank, const void* geom,
                          float* local_screen_grads, uint32_t* sent_counters, float* const* peer_inboxes,
                          int32_t* const* peer_headers, void* stream);

/* Owner side of the exchange: block (float [chunk_rows][16], the accumulation block of surfels
 * [first, first + chunk_rows)) = sum of the rows the `world` senders left in `inbox` (header[s] rows from sender s). */
EGS_API int egs_fold_inbox(int32_t chunk_rows, int32_t world, int32_t first, const float* inbox, const int32_t* header,
                           float* block, void* stream);

/*
 * Instance emission, per-tile depth sort and front-to-back compositing.  `bin` must hold `cap_instances`
 * instances (see egs_workspace_sizes); if the true count is larger the lists are truncated and
 * counters.overflow is set.  All four images are fully written (tiles without surfels get zeros, like the
 * reference's zero-initialised outputs), so the caller may pass uninitialised memory.
 * `tile_mask` must be the mask the frame's egs_forward_plan was given (NULL with NULL): the plan leaves a bit-packed
 * copy of it in `img` which the emission reads.
 */
EGS_API int egs_forward_render(const egs_frame* frame, const int32_t* tile_mask, const int32_t* radii, void* geom, void* img,
                       void* bin, int64_t cap_instances, float* out_color, float* out_normal, float* out_depth,
                       float* out_opacity, egs_counters* counters_host, int32_t flags, void* stream);

/*
 * Reverse compositing walk.  Accumulates the screen-space gradient block G[P][16] (layout above) into
 * `screen_grads`, which this call first sets to zero.  Pixel gradients are [3][H][W], [3][H][W], [H][W], [H][W].
 * `cap_instances` must be the value the binning workspace was carved with in egs_forward_render.
 */
EGS_API int egs_backward_render(const egs_frame* frame, const void* geom, const void* img, const void* bin,
                        int64_t cap_instances, const float* dL_dcolor, const float* dL_dnormal, const float* dL_ddepth,
                        const float* dL_dopacity, float* screen_grads, int32_t flags, void* stream);

/*
 * Per-surfel backward for surfels [first, first + count): conic -> cov2D -> cov3D -> scale/rotation, screen-space
 * mean and depth -> mean3D, colour -> SH.  Every output row in the range is written (zeros for culled surfels),
 * so outputs may be uninitialised.  Optional outputs (may be NULL): dL_dmeans2D [P][3], dL_dcolors [P][3],
 * dL_dcov3D [P][6].  Output pointers address row 0 of the full [P][...] arrays.
 */
EGS_API int egs_backward_surfels(const egs_frame* frame, int32_t first, int32_t count, const float* means3D, const float* shs,
                         const float* colors_precomp, const float* scales, const float* rotations,
                         const int32_t* radii, const void* geom, const float* screen_grads, float* dL_dmeans3D,
                         float* dL_dopacity, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                         float* dL_dmeans2D, float* dL_dcolors, float* dL_dcov3D, void* stream);

/* present[i] = the coarse frustum test of the reference's markVisible (auxiliary.h:152-178). */
EGS_API int egs_mark_visible(int32_t num_surfels, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/*
 * Index map of the global model in the current frame (`_C.project_surfels_to_frame`): every stable, in-frustum,
 * front-facing surfel is splatted on the 3x3 pixels around its projection; per pixel the nearest surfel wins
 * (ties: smallest index -- the race-free outcome of the reference's atomicMin + plain store).
 * index_map [height][width] i32 and depth_buffer [height][width] f32 are updated in place where a nearer surfel is
 * found (callers pre-fill them with -1 / +inf).  `scratch` = height*width*8 bytes of device memory.
 * points [P][3], rotations [P][4] (raw quaternions w,x,y,z), stable_mask [P] bytes, intrinsic [4] = fx,fy,cx,cy.
 */
EGS_API int egs_project_surfels(int32_t num_surfels, int32_t height, int32_t width, const float* points,
                                const float* rotations, const uint8_t* stable_mask, const float* intrinsic,
                                const float* viewmatrix, const float* projmatrix, void* scratch, int32_t* index_map,
                                float* depth_buffer, void* stream);

/*
 * In-place fusion of surfel position / normal with the current frame (`_C.preprocess_surfels`): mutates
 * points [P][3], rotations [P][4], sigma2 [P][2]; writes inview_mask [P], surface_mask [P] (bytes).
 * frame_vmap / frame_nmap [height][width][3], frame_dmap [height][width], frame_mask [height][width] bytes,
 * frame_imap [height][width] i32 (from egs_project_surfels).  The reference kernel's unread arguments
 * (scales, colours, confidence, tic, eta, counts, stable mask, depth buffer, model maps) are not part of the ABI.
 */
EGS_API int egs_fuse_surfels(int32_t num_surfels, int32_t height, int32_t width, const float* intrinsic,
                             const float* viewmatrix, const float* projmatrix, const float* frame_vmap,
                             const float* frame_nmap, const float* frame_dmap, const uint8_t* frame_mask,
                             const int32_t* frame_imap, float* points, float* rotations, float* sigma2,
                             uint8_t* inview_mask, uint8_t* surface_mask, float fusion_dist_thres, float alpha_p,
                             float alpha_n, void* stream);

/*
 * Test hook: copies internal index artefacts out of the workspaces into caller-provided DEVICE arrays
 * (any may be NULL): point_list[cap] u32, ranges[tiles][2] u32 (zero for empty tiles), tile_indices[tiles] i32
 * (ascending non-empty tiles, rest -1), tiles_touched[P] u32, n_contrib[H*W] u32, final_T[H*W], final_D[H*W],
 * records[P][16] f32, cov3D[P][6] f32, clamped[P] u8 (bit c = channel c clamped).
 */
EGS_API int egs_debug_export(const egs_frame* frame, const void* geom, const void* img, const void* bin,
                     int64_t cap_instances, uint32_t* point_list, uint32_t* ranges, int32_t* tile_indices,
                     uint32_t* tiles_touched, uint32_t* n_contrib, float* final_T, float* final_D, float* records,
                     float* cov3D, uint8_t* clamped, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGGSPLAT_H_ */
