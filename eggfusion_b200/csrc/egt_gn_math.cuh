// egt_gn_math.cuh -- per-pixel arithmetic of the fused dense-tracker Gauss-Newton step (egt_gn.cu), written
// __host__ __device__ so that tests/hostemu can run exactly this code on the CPU against the oracle.
// Reference: /root/reference/src/core/optimizer.py projective_transform (:131-180), icp_optimization (:317-377),
// rgb_optimization (:278-315); F.grid_sample(align_corners=True) semantics of ATen's GridSampler.
#pragma once
#include <math.h>
#include "egs_common.cuh"
#include "../../include/eggtrack.h"

struct GnWarp {
    float ut, vt, dt;   // normalised warped coordinates and inverse depth ratio
    float gx, gy;       // grid_sample coordinates in [-1, 1]
    float ix, iy;       // unnormalised source-pixel coordinates (align_corners=True)
    bool base;          // inside the ICP bound 0.98 (the wider one; NaN fails it) and model mask set
};

// projective_transform for pixel p of the level; T = row-major 4x4
EGS_HD void egt_gn_warp(const egt_level& lv, const float* T, long long p, bool in_img, GnWarp& w) {
    const int W = lv.width, H = lv.height;
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    const float us = ((float)x - lv.cx) / lv.fx, vs = ((float)y - lv.cy) / lv.fy, ds = lv.model_disp[p];
    float ut = T[0] * us + T[1] * vs + T[2] + T[3] * ds;
    float vt = T[4] * us + T[5] * vs + T[6] + T[7] * ds;
    const float zt = T[8] * us + T[9] * vs + T[10] + T[11] * ds;
    float dt = T[12] * us + T[13] * vs + T[14] + T[15] * ds;
    ut = ut / zt; vt = vt / zt; dt = dt / zt;
    w.ut = ut; w.vt = vt; w.dt = dt;
    w.gx = 2.f * (lv.fx * ut + lv.cx) / (float)(W - 1) - 1.f;
    w.gy = 2.f * (lv.fy * vt + lv.cy) / (float)(H - 1) - 1.f;
    w.base = in_img && w.gx > -0.98f && w.gx < 0.98f && w.gy > -0.98f && w.gy < 0.98f && lv.model_mask[p] != 0;
    // F.grid_sample(align_corners=True): [-1, 1] -> [0, size - 1]
    w.ix = (w.gx + 1.f) * 0.5f * (float)(W - 1);
    w.iy = (w.gy + 1.f) * 0.5f * (float)(H - 1);
}

// icp_optimization: point-to-plane row of pixel p.  Returns the pixel's weight (all masks and thresholds).
EGS_HD bool egt_gn_icp_row(const egt_level& lv, const float* T, long long p, const GnWarp& w, float sine_thres,
                           float dist_thres, float (&J)[6], float& r) {
    const int W = lv.width, H = lv.height;
    if (!(w.base && lv.frame_mask[p] != 0)) return false;   // mask_curr at the SAME pixel, not warped (as the reference)
    const float* vp = lv.model_vertex + 3 * p;
    const float* np_ = lv.model_normal + 3 * p;
    const float v0 = vp[0], v1 = vp[1], v2 = vp[2], n0 = np_[0], n1 = np_[1], n2 = np_[2];
    const float pv0 = T[0] * v0 + T[1] * v1 + T[2] * v2 + T[3];
    const float pv1 = T[4] * v0 + T[5] * v1 + T[6] * v2 + T[7];
    const float pv2 = T[8] * v0 + T[9] * v1 + T[10] * v2 + T[11];
    const float pn0 = T[0] * n0 + T[1] * n1 + T[2] * n2;
    const float pn1 = T[4] * n0 + T[5] * n1 + T[6] * n2;
    const float pn2 = T[8] * n0 + T[9] * n1 + T[10] * n2;
    // nearest, padding border, align_corners: clip then round half to even
    const int sx = (int)nearbyintf(fminf(fmaxf(w.ix, 0.f), (float)(W - 1)));
    const int sy = (int)nearbyintf(fminf(fmaxf(w.iy, 0.f), (float)(H - 1)));
    const long long q = (long long)sy * W + sx;
    const float* vc = lv.frame_vertex + 3 * q;
    const float* nc = lv.frame_normal + 3 * q;
    const float c0 = nc[0], c1 = nc[1], c2 = nc[2];
    const float d0 = vc[0] - pv0, d1 = vc[1] - pv1, d2 = vc[2] - pv2;
    const float x0 = c1 * pn2 - c2 * pn1, x1 = c2 * pn0 - c0 * pn2, x2 = c0 * pn1 - c1 * pn0;   // cross(ncurr, nprev)
    const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2), sine = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
    const bool nan_ok = x0 == x0 && x1 == x1 && x2 == x2;
    r = c0 * d0 + c1 * d1 + c2 * d2;
    J[0] = c0; J[1] = c1; J[2] = c2;                                  // J = [ncurr, cross(vprev, ncurr)]
    J[3] = pv1 * c2 - pv2 * c1; J[4] = pv2 * c0 - pv0 * c2; J[5] = pv0 * c1 - pv1 * c0;
    return nan_ok && pv2 > 0.f && sine < sine_thres && dist < dist_thres;
}

// rgb_optimization: photometric row of pixel p
EGS_HD bool egt_gn_rgb_row(const egt_level& lv, long long p, const GnWarp& w, float (&J)[6], float& r) {
    const int W = lv.width, H = lv.height;
    if (!(w.base && w.gx > -0.90f && w.gx < 0.90f && w.gy > -0.90f && w.gy < 0.90f && lv.frame_grad[3 * p + 2] > 1.f))
        return false;
    // mask_curr: nearest, padding zeros
    const int mx = (int)nearbyintf(w.ix), my = (int)nearbyintf(w.iy);
    if (!(mx >= 0 && mx < W && my >= 0 && my < H && lv.frame_mask[(long long)my * W + mx] != 0)) return false;
    // bilinear, padding zeros, align_corners
    const float fx0 = floorf(w.ix), fy0 = floorf(w.iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float tx = w.ix - fx0, ty = w.iy - fy0;
    const float wgt[4] = {(1.f - tx) * (1.f - ty), tx * (1.f - ty), (1.f - tx) * ty, tx * ty};
    float sI = 0.f, sgx = 0.f, sgy = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
        if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
            const long long q = (long long)yy * W + xx;
            sI = fmaf(lv.frame_intensity[q], wgt[k], sI);
            sgx = fmaf(lv.frame_grad[3 * q], wgt[k], sgx);
            sgy = fmaf(lv.frame_grad[3 * q + 1], wgt[k], sgy);
        }
    }
    r = lv.model_intensity[p] - sI;
    // J = Ji (1x2) @ Jc (2x6), Jc rows as in projective_transform
    const float a = sgx * lv.fx, b = sgy * lv.fy, ut = w.ut, vt = w.vt, dt = w.dt;
    J[0] = a * dt; J[1] = b * dt; J[2] = -(a * ut + b * vt) * dt;
    J[3] = -a * ut * vt - b * (1.f + vt * vt); J[4] = a * (1.f + ut * ut) + b * ut * vt; J[5] = -a * vt + b * ut;
    return true;
}
