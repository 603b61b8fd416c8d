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

/* x = solve((A + lm I) x = b) for one dense n x n system, n <= 16, entirely on the device (A row- or column-major:
 * the tracker's A is symmetric).  Singular systems yield zeros. */
EGS_API int egt_solve_block(const float* A, const float* b, float lm, float* x, int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGGTRACK_H_ */
