/*
 * eggmap.h -- C ABI of the mapping-iteration glue around the rasterizer in libeggsplat.so (sm_100a).
 *
 * SURVEY.md 8(f) row N1.  One iteration of the reference's map optimisation
 * (/root/reference/src/core/mapper.py:336-368, Mapper.frame_batch_optimization) is
 *     total_params (activations)  ->  render  ->  compute_loss  ->  loss.backward()  ->  Adam.step()
 * Everything except `render` is PyTorch glue there: ~60 elementwise / reduction launches, boolean-mask indexing
 * (a host sync each), 10 check_nan passes and a `loss.item()` per iteration.  Here it is four launches:
 *
 *   egm_loss_seed     Mapper.compute_loss, image terms (mapper.py:381-426,437-438): masked-mean colour L1, depth L1 and
 *                     normal cosine distance, PLUS the gradient seeds dL/d{color,depth,normal} that
 *                     loss.backward() would hand to the rasterizer (torch autograd of the same expressions:
 *                     abs/mean/index backward, F.cosine_similarity backward incl. its eps-clamped norms, clamp mask)
 *   egm_adam_step     (a) backward of the activations of GaussianSurfels (gaussian_surfels.py:345-425 get_opacity =
 *                     sigmoid, get_scaling = exp, get_rotation = F.normalize, then nan_to_num in
 *                     Mapper.total_params, mapper.py:565-585), (b) the regulariser of compute_loss
 *                     (mapper.py:427-435: ||pos0 - xyz||_F + reg_weight_n * mean|1 - cos(normal0, get_normal)|) and
 *                     its gradient through get_normal / build_rotation (gaussian_surfels.py:381-393,
 *                     core/utils.py:69-92), (c) torch.optim.Adam.step() for the six parameter groups of
 *                     GaussianSurfels.parametrize (gaussian_surfels.py:134-150; defaults betas (0.9, 0.999),
 *                     eps 1e-8, no weight decay / amsgrad), (d) the activations for the NEXT iteration
 *   egm_activate      the activations alone + get_normal (start of an optimisation window: pos0 / normal0 anchors)
 *   egm_loss_total    mapper.py:437-438: the weighted total from the accumulated partial sums (no host sync)
 *
 * Plain device pointers, sizes and a stream; nothing allocates or synchronises.  Return value as in eggsplat.h.
 */
#ifndef EGGMAP_H_
#define EGGMAP_H_

#include <stdint.h>
#include "eggsplat.h"

#ifdef __cplusplus
extern "C" {
#endif

/* partial sums of the image terms, device double[EGM_TERMS] */
#define EGM_TERMS 8
#define EGM_T_COUNT 0   /* pixels with rgb_mask & geo_mask */
#define EGM_T_COLOR 1   /* sum |ref - est| over masked pixels and 3 channels */
#define EGM_T_DEPTH 2   /* sum |ref - est| over masked pixels */
#define EGM_T_NORMAL 3  /* sum |1 - clamp(cos)| over masked pixels */
#define EGM_T_NAN 4     /* number of NaN values seen in the six images (the reference's check_nan passes) */

/* regulariser state, device double[EGM_REG]: slot (step & 1) holds sum (pos0 - xyz)^2 of the parameters the step
 * starts from (the caller zero-initialises slot 1 before step 1: pos0 is a copy of xyz); egm_adam_step writes the
 * same sum for the updated parameters into slot ((step + 1) & 1) and sum |1 - clamp(cos)| into slot 2. */
#define EGM_REG 4

typedef struct egm_adam {
    double beta1, beta2, eps;   /* python floats in torch: 1 - beta and the bias corrections are formed in double */
    float lr_xyz, lr_f_dc, lr_f_rest, lr_opacity, lr_scaling, lr_rotation; /* parametrize(): feature_lr, feature_lr/20 */
    int32_t step;               /* 1-based index of this update (torch's state["step"] after the increment) */
    float reg_weight;           /* cfg.Mapping.reg_weight   (0: regulariser off, pos0/normal0/reg may be NULL) */
    float reg_weight_n;         /* cfg.Mapping.reg_weight_n */
} egm_adam;

/* Image terms of Mapper.compute_loss + their gradient seeds.
 * est_*: rasterizer outputs, channel-major [3,H,W] / [1,H,W] / [3,H,W].  ref_*: frame maps, pixel-major [H,W,3] /
 * [H,W,1] / [H,W,3]; ref_depth / ref_normal may be NULL (term skipped, like `is not None` in the reference).
 * rgb_mask, geo_mask: bool [H,W]; geo_mask may be NULL.  Seeds are written channel-major for every pixel (zero
 * outside the mask), scaled by the term's weight / element count, ready for egs_backward_render. */
EGS_API int egm_loss_seed(int32_t height, int32_t width, const float* est_color, const float* est_depth,
                          const float* est_normal, const float* ref_color, const float* ref_depth,
                          const float* ref_normal, const uint8_t* rgb_mask, const uint8_t* geo_mask,
                          float color_weight, float depth_weight, float normal_weight, float* dL_dcolor,
                          float* dL_ddepth, float* dL_dnormal, double* terms, void* stream);

/* egm_loss_seed for one rank of a tile-sharded frame: `tile_mask` ([tiles_y][tiles_x] int32, the mask the
 * rank renders with; NULL = all tiles) restricts the partial sums and the seeds to the pixels of the rank's tiles, while
 * the means' denominator stays the number of masked pixels of the WHOLE frame -- so the ranks' terms[1..4] add up
 * (all-reduce) to the single-GPU values and their seeds are the single-GPU seeds on disjoint pixel sets. */
EGS_API int egm_loss_seed_tiles(int32_t height, int32_t width, const float* est_color, const float* est_depth,
                                const float* est_normal, const float* ref_color, const float* ref_depth,
                                const float* ref_normal, const uint8_t* rgb_mask, const uint8_t* geo_mask,
                                const int32_t* tile_mask, float color_weight, float depth_weight, float normal_weight,
                                float* dL_dcolor, float* dL_ddepth, float* dL_dnormal, double* terms, void* stream);

/* One fused optimiser step over P surfels with sh_coeffs SH coefficients.
 * raw parameters (updated in place): xyz[P,3], shs[P,M,3] (row 0 = _features_dc, rows 1.. = _features_rest; identity
 * activation), opacity_raw[P,1] (logit), scaling_raw[P,3] (log), rotation_raw[P,4].
 * d_*: gradients w.r.t. the ACTIVATED parameters as egs_backward_surfels writes them.
 * m_* / v_*: Adam exp_avg / exp_avg_sq, shapes of the raw parameters.
 * pos0[P,3], normal0[P,3]: regulariser anchors.  opacity/scales/rotations: activated outputs for the next forward. */
EGS_API int egm_adam_step(int32_t P, int32_t sh_coeffs, const egm_adam* hyper, float* xyz, float* shs,
                          float* opacity_raw, float* scaling_raw, float* rotation_raw, const float* d_xyz,
                          const float* d_shs, const float* d_opacity, const float* d_scales, const float* d_rotations,
                          float* m_xyz, float* v_xyz, float* m_shs, float* v_shs, float* m_opacity, float* v_opacity,
                          float* m_scaling, float* v_scaling, float* m_rotation, float* v_rotation, const float* pos0,
                          const float* normal0, double* reg, float* opacity, float* scales, float* rotations,
                          void* stream);

/* Per-surfel backward of the rasterizer (egs_backward_surfels, include/eggsplat.h) for surfels [first, first + count)
 * with the Adam update of the SH block (the f_dc / f_rest groups of GaussianSurfels.parametrize,
 * gaussian_surfels.py:134-150; torch.optim.Adam as in egm_adam_step) applied in the same kernel: dL/dSH is consumed
 * from shared memory and never reaches HBM (saves 384 B/surfel of traffic and one launch per mapping iteration).
 * `shs` [P][16][3] is read by the backward and updated in place, m_shs / v_shs are its Adam state (pointers to row 0:
 * the call offsets them by `first` itself); the other gradients are written like egs_backward_surfels does.
 * Follow it with egm_adam_step(..., sh_coeffs = 0, ...) for the remaining four groups, SAME hyper->step.
 * Only for 16 SH coefficients and 16-byte aligned arrays: EGS_E_UNSUPPORTED otherwise (use the two separate calls). */
EGS_API int egm_backward_surfels_adam(const egs_frame* frame, int32_t first, int32_t count, const float* means3D,
                                      float* shs, const float* scales, const float* rotations, const int32_t* radii,
                                      const void* geom, const float* screen_grads, float* dL_dmeans3D,
                                      float* dL_dopacity, float* dL_dscales, float* dL_drotations,
                                      const egm_adam* hyper, float* m_shs, float* v_shs, void* stream);

/* opacity = sigmoid(opacity_raw), scales = exp(scaling_raw), rotations = nan_to_num(normalize(rotation_raw), 1),
 * normals (may be NULL) = GaussianSurfels.get_normal. */
EGS_API int egm_activate(int32_t P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                         float* opacity, float* scales, float* rotations, float* normals, void* stream);

/* out[5] (device float) = total, color_loss, depth_loss, normal_loss, reg_loss of the iteration whose image terms are
 * in `terms` and whose regulariser sums are in `reg` (step = the egm_adam_step index that produced them; reg may be
 * NULL).  An empty mask gives color_loss = NaN (mean of an empty tensor) and zero depth / normal terms, as in the
 * reference. */
EGS_API int egm_loss_total(const double* terms, const double* reg, int32_t step, int32_t P, float color_weight,
                           float depth_weight, float normal_weight, float reg_weight, float reg_weight_n,
                           int32_t have_depth, int32_t have_normal, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGGMAP_H_ */
