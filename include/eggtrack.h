/*
 * eggtrack.h -- C ABI of the dense-tracking image utilities in libeggsplat.so (sm_100a).
 *
 * Replaces the live functions of the reference's torch extension `cuda_tracking_ext`
 * (/root/reference/src/utils/cuda/src/tracking.cu, "TRK"; bound at TRK:952-962, wrapped by
 * /root/reference/src/utils/cuda/__init__.py).  Plain device pointers + sizes + a stream; nothing allocates,
 * synchronises the device or uploads __constant__ tables (the reference does all three on every call).
 * Images are row-major [height][width][channels] float32.  Return value as in eggsplat.h.
 *
 *   egt_bilateral_filter      bilateral_filter_kernel        TRK:777-848   (live: frame.py:84,132)
 *   egt_gaussian_filter       gaussian_filter_kernel         TRK:705-774   (imported by frame.py:17, never called)
 *   egt_gaussian_downsample   gaussian_downsample_kernel     TRK:533-599   (live: frame.py:77-93)
 *   egt_compute_gradients     gradient_kernel                TRK:853-926   (live: frame.py:72,97, system.py:92)
 *   egt_vertex_normal_map     compute_vertex/normal_map_kernel TRK:602-702 (live: frame.py:42, mapper.py:260)
 *   egt_solve_block           solveBlock (CPU Eigen QR)      TRK:929-950   (live: tracker.py:238)
 *   egt_ingest_frame          Frame.__init__ + PyraImageCUDA (src/utils/frame.py:32-146) as one fused chain    
 *   egt_gn_accumulate         Tracker.tracking_optimization, first half (src/core/tracker.py:194-227): the PyTorch
 *                             functions projective_transform (src/core/optimizer.py:131-180), icp_optimization
 *                             (:317-377) and rgb_optimization (:278-315) fused into one pass        
 *   egt_track_pyramid         the whole coarse-to-fine loop of Tracker.tracking_frame (tracker.py:153-164) in one call
 *   egt_gn_solve_update       second half (tracker.py:229-251): combine, solve_block, convergence test, and
 *                             update_transform (optimizer.py:426-441) applied to the pose on the device
 * The three remaining exports of the reference (projective_transform / rgb_optimization / icp_optimization) are
 * dead or non-functional there and are not part of this ABI.
 */
#ifndef EGGTRACK_H_
#define EGGTRACK_H_

#include <stdint.h>
#include "eggsplat.h"

#ifdef __cplusplus
extern "C" {
#endif

/* out = bilateral(in): window x window taps, weight exp(-d2/(2 sigma_s^2) - dc2/(2 sigma_c^2)), out-of-image taps skipped. */
EGS_API int egt_bilateral_filter(const float* in, float* out, int32_t width, int32_t height, int32_t window,
                                 float sigma_color, float sigma_space, void* stream);

/* out = gaussian blur of an image with `channels` <= 4 interleaved channels, out-of-image taps skipped. */
EGS_API int egt_gaussian_filter(const float* in, float* out, int32_t width, int32_t height, int32_t channels,
                                int32_t window, float sigma_space, void* stream);

/* out [height/2][width/2][channels] = 5x5 binomial (1 4 6 4 1)^2 at stride 2, normalised by the in-image weight. */
EGS_API int egt_gaussian_downsample(const float* in, float* out, int32_t width, int32_t height, int32_t channels,
                                    void* stream);

/* 3x3 Scharr-like derivative pair with the reference's coefficients and (reversed) tap order. */
EGS_API int egt_compute_gradients(const float* in, float* grad_x, float* grad_y, int32_t width, int32_t height,
                                  void* stream);

/* vertex_map [h][w][3] = back-projection of depth; normal_map = normalize(cross(v(x,y+1)-v, v(x+1,y)-v)), NaN -> 0. */
EGS_API int egt_vertex_normal_map(const float* depth, float fx, float fy, float cx, float cy, float* vertex_map,
                                  float* normal_map, int32_t width, int32_t height, void* stream);

/* x = (A' + lm I).colPivHouseholderQr().solve(b) for one dense n x n system, n <= 16, entirely on the device: the
 * reference's algorithm (Eigen's column-pivoted Householder QR, fp32) on the reference's view of the buffer (A' = the
 * n x n buffer read COLUMN-major, i.e. the transpose of torch's row-major matrix; the tracker's A is symmetric).
 * Rank-deficient systems get Eigen's basic solution (zeros for the dropped pivots). */
EGS_API int egt_solve_block(const float* A, const float* b, float lm, float* x, int32_t n, void* stream);

/* One pyramid level of the model (rendered, "prev"/frame1) and of the incoming frame ("curr"/frame2), PyraImageCUDA
 * layout (src/utils/frame.py:20-99): row-major [height][width][C] float32, masks bool [height][width]. */
typedef struct egt_level {
    int32_t width, height;
    float fx, fy, cx, cy;            /* intrinsic_pyramid[level] */
    const float* model_disp;         /* disp_pyramid      [h][w][1] */
    const float* model_vertex;       /* vertex_pyramid    [h][w][3] */
    const float* model_normal;       /* normal_pyramid    [h][w][3] */
    const uint8_t* model_mask;       /* mask_pyramid      [h][w][1] */
    const float* model_intensity;    /* intensity_pyramid [h][w][1] (rgb term only) */
    const float* frame_vertex;
    const float* frame_normal;
    const uint8_t* frame_mask;
    const float* frame_intensity;    /* (rgb term only) */
    const float* frame_grad;         /* grad_pyramid [h][w][3] = d/dx, d/dy, magnitude (rgb term only) */
} egt_level;

/* sums (device double[EGT_GN_SUMS]): [0,21) upper triangle of J^T J of the ICP term (row-major: 00 01 .. 05 11 ..),
 * [21,27) its J^T r, [27] its valid-pixel count; [28,56) the same for the photometric term. */
#define EGT_GN_SUMS 56

/* transform: device float[16], row-major 4x4 (the current dense_delta).  Zeroes `sums`, then accumulates. */
EGS_API int egt_gn_accumulate(const egt_level* level, const float* transform, float angle_thres_deg, float dist_thres,
                              int32_t use_rgb, double* sums, void* stream);

/* A = A_icp + rgb_weight A_rgb, b likewise; dx = solve((A + lm I) dx = b); converged = ||b|| / max(1, sqrt(count)) <
 * residual_thres && ||dx|| < dx_thres; transform <- update_transform(transform, dx) in place.
 * dx_out: float[6] or NULL; system_out: float[42] = A (row-major 6x6) then b, or NULL;
 * status: int32[4] or NULL: [0] |= converged (caller zeroes it per frame), [1] converged, [2] icp count, [3] rgb count. */
EGS_API int egt_gn_solve_update(const double* sums, float rgb_weight, float lm, float residual_thres, float dx_thres,
                                float* transform, float* dx_out, float* system_out, int32_t* status, void* stream);

/* The dense loop of Tracker.tracking_frame (tracker.py:153-164) in ONE call: for l in 0..nlevel-1, iters[l] Gauss-Newton
 * steps on pyramid level nlevel-1-l (levels[] is indexed by pyramid level), each = egt_gn_accumulate + egt_gn_solve_update
 * on `transform` (device, in/out).  status[0] = any step converged.  2 launches per step, no host involvement. */
EGS_API int egt_track_pyramid(const egt_level* levels, int32_t nlevel, const int32_t* iters, float angle_thres_deg,
                              float dist_thres, int32_t use_rgb, float rgb_weight, float lm, float residual_thres,
                              float dx_thres, float* transform, double* sums, float* dx_out, float* system_out,
                              int32_t* status, void* stream);

/* One level of a frame's pyramids as the tracker consumes them (PyraImageCUDA, src/utils/frame.py:22-99): row-major
 * [height][width][C] float32 device arrays, caller-owned.  maskf is the float mask chain the reference keeps
 * downsampling (frame.py:86); mask the bool map it derives per level (frame.py:69,88). */
typedef struct egt_pyramid_level {
    int32_t width, height;
    float* depth;     /* [h][w]    bilateral-filtered depth (level 0: Frame.depth, frame.py:132; level l: frame.py:83-84) */
    float* disp;      /* [h][w]    1 / (depth + 1e-6)                                  disp_pyramid */
    uint8_t* mask;    /* [h][w]    (maskf > 0.9) & (depth > 0.1)                       mask_pyramid */
    float* maskf;     /* [h][w]    float mask at this level */
    float* vertex;    /* [h][w][3]                                                     vertex_pyramid */
    float* normal;    /* [h][w][3]                                                     normal_pyramid */
    float* gray;      /* [h][w]                                                        intensity_pyramid */
    float* grad;      /* [h][w][3] d/dx, d/dy, sqrt(dx^2 + dy^2 + 1e-6)                grad_pyramid */
} egt_pyramid_level;

/*
 * Frame ingest: everything Frame.__init__ + PyraImageCUDA compute on the device for one RGB-D
 * frame (src/utils/frame.py:112-146, :32-99) as `nlevel` stream-ordered launches -- the 13x13 bilateral of the raw
 * depth, vertex / normal map, grey image, derivatives and, per further level, the 5x5 stride-2 downsamples of grey /
 * depth / mask / vertex / normal with the bilateral of the downsampled depth and the re-normalised normals.
 * color [height][width][3] in 0..1, depth_raw [height][width] in metres (unfiltered), mask [height][width] float;
 * levels[l] must be sized width >> l by height >> l.  Level l's intrinsics are (fx, fy, cx, cy) / 2^l ... the reference
 * divides the PREVIOUS level's by 2^l (frame.py:80-81: level 2 = level 0 / 8); that list is host data and stays in
 * the Python mirror (eggfusion_b200.tracking.ingest_frame).
 */
EGS_API int egt_ingest_frame(const float* color, const float* depth_raw, const float* mask, int32_t width, int32_t height,
                             float fx, float fy, float cx, float cy, float sigma_color, float sigma_space, int32_t nlevel,
                             const egt_pyramid_level* levels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGGTRACK_H_ */
