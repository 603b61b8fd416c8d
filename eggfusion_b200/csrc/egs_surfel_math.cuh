// egs_surfel_math.cuh -- per-surfel forward projection and backward, as host/device inline functions.
//
// Forward: what the reference's preprocessCUDA computes per surfel
//   (DGS/cuda_rasterizer/forward.cu:158-301 with helpers auxiliary.h:42-57,140-149,180-281, forward.cu:20-155).
// Backward: computeCov2DCUDA + preprocessCUDA(bwd) + computeColorFromSH(bwd) + computeCov3D(bwd)
//   (DGS/cuda_rasterizer/backward.cu:20-416) fused into one pass, including the reference's deliberate
//   deviations from the true derivative (SURVEY.md 8 a-bis).
// The functions are __host__ __device__ so tests/hostemu can execute the very same arithmetic on the CPU.
#pragma once
#include "egs_common.cuh"

#define EGS_SH_C0 0.28209479177387814f
#define EGS_SH_C1 0.4886025119029199f
#define EGS_SH_C2_0 1.0925484305920792f
#define EGS_SH_C2_1 -1.0925484305920792f
#define EGS_SH_C2_2 0.31539156525252005f
#define EGS_SH_C2_3 -1.0925484305920792f
#define EGS_SH_C2_4 0.5462742152960396f
#define EGS_SH_C3_0 -0.5900435899266435f
#define EGS_SH_C3_1 2.890611442640554f
#define EGS_SH_C3_2 -0.4570457994644658f
#define EGS_SH_C3_3 0.3731763325901154f
#define EGS_SH_C3_4 -0.4570457994644658f
#define EGS_SH_C3_5 1.445305721320277f
#define EGS_SH_C3_6 -0.5900435899266435f

struct Mat3 { float m[3][3]; };

// One output row of the row-vector affine transform p * M (M given as 16 floats, element [4*c + r]).
EGS_HD float xf_affine(const float* M, int r, float x, float y, float z) {
    return f_add(f_dot3(M[r], x, M[4 + r], y, M[8 + r], z), M[12 + r]);
}
EGS_HD float xf_linear(const float* M, int r, float x, float y, float z) {
    return f_dot3(M[r], x, M[4 + r], y, M[8 + r], z);
}

// Rotation matrix of quaternion (w,x,y,z) exactly as the reference evaluates it; R.m[i][j] is the usual
// (row i, column j) entry, so column 2 is the surfel normal and columns 0/1 its tangent axes.
EGS_HD Mat3 quat_to_rot(float r, float x, float y, float z) {
    Mat3 R;
    R.m[0][0] = f_fma(-2.f, f_fma(y, y, f_mul(z, z)), 1.f);
    R.m[0][1] = f_mul(2.f, f_fma(x, y, -f_mul(r, z)));
    R.m[0][2] = f_mul(2.f, f_fma(x, z, f_mul(r, y)));
    R.m[1][0] = f_mul(2.f, f_fma(x, y, f_mul(r, z)));
    R.m[1][1] = f_fma(-2.f, f_fma(x, x, f_mul(z, z)), 1.f);
    R.m[1][2] = f_mul(2.f, f_fma(y, z, -f_mul(r, x)));
    R.m[2][0] = f_mul(2.f, f_fma(x, z, -f_mul(r, y)));
    R.m[2][1] = f_mul(2.f, f_fma(y, z, f_mul(r, x)));
    R.m[2][2] = f_fma(-2.f, f_fma(x, x, f_mul(y, y)), 1.f);
    return R;
}

// EWA projection matrix T = W * J restricted to its two non-zero columns; t is the (clamped) view-space mean.
struct EwaT {
    float T0[3], T1[3]; // T0[r] = T[0][r], T1[r] = T[1][r] in the reference's glm indexing
    float tx, ty, tz;   // clamped t
    float txtz, tytz;   // unclamped ratios
};
EGS_HD EwaT ewa_T(const FrameConst& fc, float vx, float vy, float vz) {
    EwaT e;
    const float limx = f_mul(1.3f, fc.tanfovx), limy = f_mul(1.3f, fc.tanfovy);
    e.txtz = f_div(vx, vz);
    e.tytz = f_div(vy, vz);
    e.tx = f_mul(fminf(limx, fmaxf(-limx, e.txtz)), vz);
    e.ty = f_mul(fminf(limy, fmaxf(-limy, e.tytz)), vz);
    e.tz = vz;
    const float zz = f_mul(vz, vz);
    const float J00 = f_div(fc.fx, vz), J02 = f_div(-f_mul(fc.fx, e.tx), zz);
    const float J11 = f_div(fc.fy, vz), J12 = f_div(-f_mul(fc.fy, e.ty), zz);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const float W0 = fc.view[4 * r], W1 = fc.view[4 * r + 1], W2 = fc.view[4 * r + 2];
        e.T0[r] = f_fma(W2, J02, f_mul(W0, J00));
        e.T1[r] = f_fma(W2, J12, f_mul(W1, J11));
    }
    return e;
}
// cov2D = T^T Vrk T with the +0.3 low-pass on the diagonal.
EGS_HD void ewa_cov2d(const EwaT& e, const float* c6, float& a, float& b, float& c) {
    const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float A0[3], A1[3]; // A0[k] = sum_j T0[j] V[k][j]
#pragma unroll
    for (int k = 0; k < 3; k++) {
        A0[k] = f_dot3(e.T0[0], V[k][0], e.T0[1], V[k][1], e.T0[2], V[k][2]);
        A1[k] = f_dot3(e.T1[0], V[k][0], e.T1[1], V[k][1], e.T1[2], V[k][2]);
    }
    a = f_add(f_dot3(A0[0], e.T0[0], A0[1], e.T0[1], A0[2], e.T0[2]), 0.3f);
    b = f_dot3(A1[0], e.T0[0], A1[1], e.T0[1], A1[2], e.T0[2]);
    c = f_add(f_dot3(A1[0], e.T1[0], A1[1], e.T1[1], A1[2], e.T1[2]), 0.3f);
}

// SH -> RGB (+0.5, clamp at 0).  `sh` points at this surfel's [M][3] block; returns the clamp bitmask.
EGS_HD uint32_t sh_eval(int deg, const float* sh, float dx, float dy, float dz, float rgb[3]) {
    const float len = f_sqrt(f_dot3(dx, dx, dy, dy, dz, dz));
    const float x = f_div(dx, len), y = f_div(dy, len), z = f_div(dz, len);
    uint32_t clamp_bits = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
#define S_(k) sh[3 * (k) + ch]
        float res = f_mul(EGS_SH_C0, S_(0));
        if (deg > 0) {
            res = f_fma(-f_mul(EGS_SH_C1, y), S_(1), res);
            res = f_fma(f_mul(EGS_SH_C1, z), S_(2), res);
            res = f_fma(-f_mul(EGS_SH_C1, x), S_(3), res);
            if (deg > 1) {
                const float xx = f_mul(x, x), yy = f_mul(y, y), zz = f_mul(z, z);
                const float xy = f_mul(x, y), yz = f_mul(y, z), xz = f_mul(x, z);
                res = f_fma(f_mul(EGS_SH_C2_0, xy), S_(4), res);
                res = f_fma(f_mul(EGS_SH_C2_1, yz), S_(5), res);
                res = f_fma(f_mul(EGS_SH_C2_2, f_sub(f_fma(2.0f, zz, -xx), yy)), S_(6), res);
                res = f_fma(f_mul(EGS_SH_C2_3, xz), S_(7), res);
                res = f_fma(f_mul(EGS_SH_C2_4, f_sub(xx, yy)), S_(8), res);
                if (deg > 2) {
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_0, y), f_fma(3.0f, xx, -yy)), S_(9), res);
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_1, xy), z), S_(10), res);
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_2, y), f_sub(f_fma(4.0f, zz, -xx), yy)), S_(11), res);
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_3, z), f_fma(-3.0f, yy, f_fma(2.0f, zz, -f_mul(3.0f, xx)))), S_(12), res);
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_4, x), f_sub(f_fma(4.0f, zz, -xx), yy)), S_(13), res);
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_5, z), f_sub(xx, yy)), S_(14), res);
                    res = f_fma(f_mul(f_mul(EGS_SH_C3_6, x), f_fma(-3.0f, yy, xx)), S_(15), res);
                }
            }
        }
#undef S_
        res = f_add(res, 0.5f);
        if (res < 0.f) clamp_bits |= 1u << ch;
        rgb[ch] = fmaxf(res, 0.f);
    }
    return clamp_bits;
}

// Result of the forward per-surfel stage.
struct SurfelFwd {
    int radius;      // 0 => culled
    int active;      // inside the frustum (set before the back-face test, like the reference)
    int x0, y0, x1, y1; // tile rectangle (valid when radius > 0)
    float cov3D[6];
    uint32_t clamped;
    SplatRecord rec;
};

// Conservative half extents (1/8 px, u16 each) of the region where alpha can reach 1/255.
EGS_HD uint32_t pack_extent(float cov_xx, float cov_yy, float opacity) {
    const float s = 255.f * opacity;
    if (!(s > 1.0f)) return 0u; // alpha = min(.99, o*exp(power<=0)) < 1/255 everywhere
    const float tau = 2.f * logf(s) + 0.02f;
    const float hx = sqrtf(tau * cov_xx) * 1.002f + 0.02f;
    const float hy = sqrtf(tau * cov_yy) * 1.002f + 0.02f;
    const float qx = fminf(65535.f, ceilf(hx * 8.f) + 1.f), qy = fminf(65535.f, ceilf(hy * 8.f) + 1.f);
    return (uint32_t)qx | ((uint32_t)qy << 16);
}

// Colour of a visible surfel: color_src = its SH block [M][3] (use_sh; may live in registers) or an RGB triple.
EGS_HD void surfel_color(const FrameConst& fc, const float* mean, const float* color_src, bool use_sh, SurfelFwd& o) {
    float rgb[3];
    if (use_sh) {
        o.clamped = sh_eval(fc.D, color_src, f_sub(mean[0], fc.campos[0]), f_sub(mean[1], fc.campos[1]),
                            f_sub(mean[2], fc.campos[2]), rgb);
    } else {
        rgb[0] = color_src[0]; rgb[1] = color_src[1]; rgb[2] = color_src[2];
    }
    o.rec.r = rgb[0]; o.rec.g = rgb[1]; o.rec.b = rgb[2];
}

// Geometry of one surfel (mean / scale / rot: its rows).  Leaves o.radius == 0 when culled; the colour of a
// visible surfel is filled in afterwards by surfel_color().
EGS_HD void surfel_forward(const FrameConst& fc, const float* mean, const float* scale, const float* rot, float opacity,
                           SurfelFwd& o) {
    o.radius = 0;
    o.active = 0;
    o.clamped = 0;
    const float px = mean[0], py = mean[1], pz = mean[2];
    const float hx = xf_affine(fc.proj, 0, px, py, pz), hy = xf_affine(fc.proj, 1, px, py, pz);
    const float hw = xf_affine(fc.proj, 3, px, py, pz);
    const float pw = f_rcp(f_add(hw, 0.0000001f));
    const float ndx = f_mul(hx, pw), ndy = f_mul(hy, pw);
    const float vx = xf_affine(fc.view, 0, px, py, pz), vy = xf_affine(fc.view, 1, px, py, pz);
    const float vz = xf_affine(fc.view, 2, px, py, pz);
    // ndc2Pix: fp32 product, a single fp64 fma, back to fp32 (auxiliary.h:42-45)
    const float ix = (float)fma((double)f_mul(ndx, (float)fc.W), 0.5, (double)fc.cx);
    const float iy = (float)fma((double)f_mul(ndy, (float)fc.H), 0.5, (double)fc.cy);
    {   // in_frustum (auxiliary.h:140-149)
        const float e = 0.05f, e1 = f_add(1.0f, e);
        const float x0 = f_mul((float)(-fc.W), e), x1 = f_mul((float)fc.W, e1);
        const float y0 = f_mul((float)(-fc.H), e), y1 = f_mul((float)fc.H, e1);
        if (vz < 0.f || ix < x0 || ix >= x1 || iy < y0 || iy >= y1) return;
    }
    o.active = 1;

    const Mat3 R = quat_to_rot(rot[0], rot[1], rot[2], rot[3]);
    float nv[3], a0[3], a1[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        nv[r] = xf_linear(fc.view, r, R.m[0][2], R.m[1][2], R.m[2][2]);
        a0[r] = xf_linear(fc.view, r, R.m[0][0], R.m[1][0], R.m[2][0]);
        a1[r] = xf_linear(fc.view, r, R.m[0][1], R.m[1][1], R.m[2][1]);
    }
    // front_facing (auxiliary.h:180-194): compared in double against -1e-5
    const float facing = f_dot3(vx, nv[0], vy, nv[1], vz, nv[2]);
    if ((double)facing > -0.00001) return;

    // local_homo (auxiliary.h:205-281)
    float J0, J1, J2, J3;
    {
        const float prx = f_div(vx, vz), pry = f_div(vy, vz);
        const float s_fix = 1000.f, inv_fix = 0.001f; // 1 / S_fix folded at compile time in the reference
        float d0[3] = {f_add(prx, inv_fix), pry, 1.f};
        float d1[3] = {prx, f_add(pry, inv_fix), 1.f};
        const float m0 = fmaxf(f_sqrt(f_dot3(d0[0], d0[0], d0[1], d0[1], d0[2], d0[2])), 0.00000001f);
        const float m1 = fmaxf(f_sqrt(f_dot3(d1[0], d1[0], d1[1], d1[1], d1[2], d1[2])), 0.00000001f);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            d0[k] = f_div(d0[k], m0);
            d1[k] = f_div(d1[k], m1);
        }
        const float prj0 = f_dot3(d0[0], nv[0], d0[1], nv[1], d0[2], nv[2]);
        const float prj1 = f_dot3(d1[0], nv[0], d1[1], nv[1], d1[2], nv[2]);
        if (fabsf(f_div(prj0, m0)) < 0.01f || fabsf(f_div(prj1, m1)) < 0.01f) return; // grazing
        const float t0 = f_div(facing, prj0), t1 = f_div(facing, prj1);
        const float pv[3] = {vx, vy, vz};
        float xu0[3], xu1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            xu0[k] = f_fma(d0[k], t0, -pv[k]);
            xu1[k] = f_fma(d1[k], t1, -pv[k]);
        }
        const float kk = f_div(f_div(f_add(fc.fx, fc.fy), 2.f), s_fix);
        J0 = f_div(f_dot3(xu0[0], a0[0], xu0[1], a0[1], xu0[2], a0[2]), kk);
        J1 = f_div(f_dot3(xu1[0], a0[0], xu1[1], a0[1], xu1[2], a0[2]), kk);
        J2 = f_div(f_dot3(xu0[0], a1[0], xu0[1], a1[1], xu0[2], a1[2]), kk);
        J3 = f_div(f_dot3(xu1[0], a1[0], xu1[1], a1[1], xu1[2], a1[2]), kk);
    }

    // computeCov3D (forward.cu:135-155): S = diag(mod sx, mod sy, 0); only two rows of M = S R are non-zero.
    {
        const float s0 = f_mul(fc.mod, scale[0]), s1 = f_mul(fc.mod, scale[1]);
        float M0[3], M1[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            M0[c] = f_mul(s0, R.m[c][0]);
            M1[c] = f_mul(s1, R.m[c][1]);
        }
        o.cov3D[0] = f_fma(M0[0], M0[0], f_mul(M1[0], M1[0]));
        o.cov3D[1] = f_fma(M0[1], M0[0], f_mul(M1[1], M1[0]));
        o.cov3D[2] = f_fma(M0[2], M0[0], f_mul(M1[2], M1[0]));
        o.cov3D[3] = f_fma(M0[1], M0[1], f_mul(M1[1], M1[1]));
        o.cov3D[4] = f_fma(M0[2], M0[1], f_mul(M1[2], M1[1]));
        o.cov3D[5] = f_fma(M0[2], M0[2], f_mul(M1[2], M1[2]));
    }

    const EwaT e = ewa_T(fc, vx, vy, vz);
    float ca, cb, cc;
    ewa_cov2d(e, o.cov3D, ca, cb, cc);
    const float det = f_fma(ca, cc, -f_mul(cb, cb));
    if (det == 0.0f) return;
    const float det_inv = f_rcp(det);
    const float mid = f_mul(0.5f, f_add(ca, cc));
    const float sq = f_sqrt(fmaxf(0.1f, f_fma(mid, mid, -det)));
    const float lam = fmaxf(f_add(mid, sq), f_sub(mid, sq));
    const int rad = (int)ceilf(f_mul(3.f, f_sqrt(lam)));
    egs_tile_rect(ix, iy, rad, fc.gx, fc.gy, o.x0, o.y0, o.x1, o.y1);
    if ((o.x1 - o.x0) * (o.y1 - o.y0) == 0) return;

    o.radius = rad;
    SplatRecord& q = o.rec;
    q.x = ix; q.y = iy;
    q.ext = pack_extent(ca, cc, opacity);
    q.opacity = opacity;
    q.cxx = f_mul(cc, det_inv); q.cxy = f_mul(-cb, det_inv); q.cyy = f_mul(ca, det_inv);
    q.depth = vz;
    // plane-depth slope: pos_dif.z of depth_differencing (auxiliary.h:283-290) is linear in the pixel offset
    q.ja = f_fma(J0, a0[2], f_mul(J2, a1[2]));
    q.jb = f_fma(J1, a0[2], f_mul(J3, a1[2]));
    q.r = 0.f; q.g = 0.f; q.b = 0.f;
    q.nx = nv[0]; q.ny = nv[1]; q.nz = nv[2];
}

// ------------------------------------------------------------------------------------------------ backward
struct SurfelBwd {
    float d_mean[3];
    float d_scale[3];
    float d_rot[4];
    float d_cov3D[6];
};

// dL/dmean of normalize(v) given dL/d(normalized v)   (auxiliary.h:108-118)
EGS_HD void dnormalize3(const float v[3], const float dv[3], float out[3]) {
    const float s2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float inv = 1.0f / sqrtf(s2 * s2 * s2);
    out[0] = ((s2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
    out[1] = (-v[0] * v[1] * dv[0] + (s2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
    out[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (s2 - v[2] * v[2]) * dv[2]) * inv;
}

// SH backward: writes d_sh[0 .. (deg+1)^2) x 3 through `store(k, ch, value)` and returns dL/d(direction).
// All reads of `sh` happen before the first `store`, so the output may alias the input.
template <class Store>
EGS_HD void sh_backward(int deg, const float* sh, const float dir_orig[3], uint32_t clamped, const float g_in[3],
                        Store store, float d_mean_add[3]) {
    const float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
    const float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
    float g[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) g[ch] = (clamped >> ch & 1u) ? 0.f : g_in[ch];
    float ddx = 0.f, ddy = 0.f, ddz = 0.f; // dL/d(dir)
#define S_(k, ch) sh[3 * (k) + (ch)]
#define OUT_(k, coef)                                 \
    {                                                 \
        const float cf = (coef);                      \
        store(k, 0, cf * g[0]);                       \
        store(k, 1, cf * g[1]);                       \
        store(k, 2, cf * g[2]);                       \
    }
    // 1. dL/d(direction): reads every coefficient.  Done first so that `store` may overwrite `sh` in place.
    float xx = 0, yy = 0, zz = 0, xy = 0, yz = 0, xz = 0;
    if (deg > 1) { xx = x * x; yy = y * y; zz = z * z; xy = x * y; yz = y * z; xz = x * z; }
    if (deg > 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float dx = -EGS_SH_C1 * S_(3, ch), dy = -EGS_SH_C1 * S_(1, ch), dz = EGS_SH_C1 * S_(2, ch);
            if (deg > 1) {
                dx += EGS_SH_C2_0 * y * S_(4, ch) + EGS_SH_C2_2 * 2.f * -x * S_(6, ch) + EGS_SH_C2_3 * z * S_(7, ch) +
                      EGS_SH_C2_4 * 2.f * x * S_(8, ch);
                dy += EGS_SH_C2_0 * x * S_(4, ch) + EGS_SH_C2_1 * z * S_(5, ch) + EGS_SH_C2_2 * 2.f * -y * S_(6, ch) +
                      EGS_SH_C2_4 * 2.f * -y * S_(8, ch);
                dz += EGS_SH_C2_1 * y * S_(5, ch) + EGS_SH_C2_2 * 2.f * 2.f * z * S_(6, ch) + EGS_SH_C2_3 * x * S_(7, ch);
                if (deg > 2) {
                    dx += (EGS_SH_C3_0 * S_(9, ch) * 3.f * 2.f * xy + EGS_SH_C3_1 * S_(10, ch) * yz +
                           EGS_SH_C3_2 * S_(11, ch) * -2.f * xy + EGS_SH_C3_3 * S_(12, ch) * -3.f * 2.f * xz +
                           EGS_SH_C3_4 * S_(13, ch) * (-3.f * xx + 4.f * zz - yy) + EGS_SH_C3_5 * S_(14, ch) * 2.f * xz +
                           EGS_SH_C3_6 * S_(15, ch) * 3.f * (xx - yy));
                    dy += (EGS_SH_C3_0 * S_(9, ch) * 3.f * (xx - yy) + EGS_SH_C3_1 * S_(10, ch) * xz +
                           EGS_SH_C3_2 * S_(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                           EGS_SH_C3_3 * S_(12, ch) * -3.f * 2.f * yz + EGS_SH_C3_4 * S_(13, ch) * -2.f * xy +
                           EGS_SH_C3_5 * S_(14, ch) * -2.f * yz + EGS_SH_C3_6 * S_(15, ch) * -3.f * 2.f * xy);
                    dz += (EGS_SH_C3_1 * S_(10, ch) * xy + EGS_SH_C3_2 * S_(11, ch) * 4.f * 2.f * yz +
                           EGS_SH_C3_3 * S_(12, ch) * 3.f * (2.f * zz - xx - yy) +
                           EGS_SH_C3_4 * S_(13, ch) * 4.f * 2.f * xz + EGS_SH_C3_5 * S_(14, ch) * (xx - yy));
                }
            }
            ddx += dx * g[ch];
            ddy += dy * g[ch];
            ddz += dz * g[ch];
        }
    }
    // 2. dL/dSH = basis * dL/dRGB
    OUT_(0, EGS_SH_C0);
    if (deg > 0) {
        OUT_(1, -EGS_SH_C1 * y);
        OUT_(2, EGS_SH_C1 * z);
        OUT_(3, -EGS_SH_C1 * x);
        if (deg > 1) {
            OUT_(4, EGS_SH_C2_0 * xy);
            OUT_(5, EGS_SH_C2_1 * yz);
            OUT_(6, EGS_SH_C2_2 * (2.f * zz - xx - yy));
            OUT_(7, EGS_SH_C2_3 * xz);
            OUT_(8, EGS_SH_C2_4 * (xx - yy));
            if (deg > 2) {
                OUT_(9, EGS_SH_C3_0 * y * (3.f * xx - yy));
                OUT_(10, EGS_SH_C3_1 * xy * z);
                OUT_(11, EGS_SH_C3_2 * y * (4.f * zz - xx - yy));
                OUT_(12, EGS_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy));
                OUT_(13, EGS_SH_C3_4 * x * (4.f * zz - xx - yy));
                OUT_(14, EGS_SH_C3_5 * z * (xx - yy));
                OUT_(15, EGS_SH_C3_6 * x * (xx - 3.f * yy));
            }
        }
    }
#undef S_
#undef OUT_
    const float dd[3] = {ddx, ddy, ddz};
    dnormalize3(dir_orig, dd, d_mean_add);
}

// g16: this surfel's row of the screen-space gradient block (layout in eggsplat.h).
// Everything except SH, whose outputs are streamed by the caller through sh_backward().
EGS_HD void surfel_backward_geom(const FrameConst& fc, const float* mean, const float* scale, const float* rot,
                                 const float* c6, const float* g16, SurfelBwd& o) {
    const float* V = fc.view;
    const float* PM = fc.proj;
    float gm[3];
    // ---- conic -> cov2D -> cov3D and the covariance path of dL/dmean (backward.cu:144-274)
    {
        const float gcx = g16[2], gcy = g16[3], gcw = g16[4];
        const float vx = xf_affine(V, 0, mean[0], mean[1], mean[2]), vy = xf_affine(V, 1, mean[0], mean[1], mean[2]);
        const float vz = xf_affine(V, 2, mean[0], mean[1], mean[2]);
        const EwaT e = ewa_T(fc, vx, vy, vz);
        const float limx = 1.3f * fc.tanfovx, limy = 1.3f * fc.tanfovy;
        const float xmul = (e.txtz < -limx || e.txtz > limx) ? 0.f : 1.f;
        const float ymul = (e.tytz < -limy || e.tytz > limy) ? 0.f : 1.f;
        float a, b, c;
        ewa_cov2d(e, c6, a, b, c);
        const float Vr[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        const float denom = a * c - b * b;
        float da = 0.f, db = 0.f, dc = 0.f;
        const float d2inv = 1.0f / ((denom * denom) + 0.0000001f);
        const float* T0 = e.T0;
        const float* T1 = e.T1;
        if (d2inv != 0.f) {
            da = d2inv * (-c * c * gcx + 2 * b * c * gcy + (denom - a * c) * gcw);
            dc = d2inv * (-a * a * gcw + 2 * a * b * gcy + (denom - a * c) * gcx);
            db = d2inv * 2 * (b * c * gcx - (denom + 2 * b * b) * gcy + a * b * gcw);
            o.d_cov3D[0] = (T0[0] * T0[0] * da + T0[0] * T1[0] * db + T1[0] * T1[0] * dc);
            o.d_cov3D[3] = (T0[1] * T0[1] * da + T0[1] * T1[1] * db + T1[1] * T1[1] * dc);
            o.d_cov3D[5] = (T0[2] * T0[2] * da + T0[2] * T1[2] * db + T1[2] * T1[2] * dc);
            o.d_cov3D[1] = 2 * T0[0] * T0[1] * da + (T0[0] * T1[1] + T0[1] * T1[0]) * db + 2 * T1[0] * T1[1] * dc;
            o.d_cov3D[2] = 2 * T0[0] * T0[2] * da + (T0[0] * T1[2] + T0[2] * T1[0]) * db + 2 * T1[0] * T1[2] * dc;
            o.d_cov3D[4] = 2 * T0[2] * T0[1] * da + (T0[1] * T1[2] + T0[2] * T1[1]) * db + 2 * T1[1] * T1[2] * dc;
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) o.d_cov3D[k] = 0.f;
        }
        float dT0[3], dT1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float tv0 = T0[0] * Vr[k][0] + T0[1] * Vr[k][1] + T0[2] * Vr[k][2];
            const float tv1 = T1[0] * Vr[k][0] + T1[1] * Vr[k][1] + T1[2] * Vr[k][2];
            dT0[k] = 2 * tv0 * da + tv1 * db;
            dT1[k] = 2 * tv1 * dc + tv0 * db;
        }
        const float dJ00 = V[0] * dT0[0] + V[4] * dT0[1] + V[8] * dT0[2];
        const float dJ02 = V[2] * dT0[0] + V[6] * dT0[1] + V[10] * dT0[2];
        const float dJ11 = V[1] * dT1[0] + V[5] * dT1[1] + V[9] * dT1[2];
        const float dJ12 = V[2] * dT1[0] + V[6] * dT1[1] + V[10] * dT1[2];
        const float tz = 1.f / e.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = xmul * -fc.fx * tz2 * dJ02;
        const float dty = ymul * -fc.fy * tz2 * dJ12;
        const float dtz = -fc.fx * tz2 * dJ00 - fc.fy * tz2 * dJ11 + (2 * fc.fx * e.tx) * tz3 * dJ02 +
                          (2 * fc.fy * e.ty) * tz3 * dJ12;
        gm[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
        gm[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
        gm[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    }
    // ---- screen-space mean and depth paths (backward.cu:385-407)
    {
        const float hw = xf_affine(PM, 3, mean[0], mean[1], mean[2]);
        const float mw = 1.0f / (hw + 0.0000001f);
        const float mul1 = (PM[0] * mean[0] + PM[4] * mean[1] + PM[8] * mean[2] + PM[12]) * mw * mw;
        const float mul2 = (PM[1] * mean[0] + PM[5] * mean[1] + PM[9] * mean[2] + PM[13]) * mw * mw;
        const float g2x = g16[0], g2y = g16[1], gd = g16[12];
        const float dmx = (PM[0] * mw - PM[3] * mul1) * g2x + (PM[1] * mw - PM[3] * mul2) * g2y;
        const float dmy = (PM[4] * mw - PM[7] * mul1) * g2x + (PM[5] * mw - PM[7] * mul2) * g2y;
        const float dmz = (PM[8] * mw - PM[11] * mul1) * g2x + (PM[9] * mw - PM[11] * mul2) * g2y;
        gm[0] += dmx + gd * V[2];
        gm[1] += dmy + gd * V[6];
        gm[2] += dmz + gd * V[10];
    }
    o.d_mean[0] = gm[0]; o.d_mean[1] = gm[1]; o.d_mean[2] = gm[2];

    // ---- cov3D -> scale / rotation, with the normal-gradient injection (backward.cu:278-353)
    {
        const float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
        const Mat3 R = quat_to_rot(r, x, y, z);
        const float s[3] = {fc.mod * scale[0], fc.mod * scale[1], fc.mod * scale[2]}; // s.z is NOT forced to 0 here
        const float* g6 = o.d_cov3D;
        const float dS[3][3] = {{g6[0], 0.5f * g6[1], 0.5f * g6[2]},
                                {0.5f * g6[1], g6[3], 0.5f * g6[4]},
                                {0.5f * g6[2], 0.5f * g6[4], g6[5]}};
        // Mg[c][rr] = s[rr] R[c][rr];  dM[c][rr] = sum_k 2 Mg[k][rr] dS[c][k];  dMt[a][b] = dM[b][a]
        float dMt[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++)
                dMt[rr][c] = 2.0f * (s[rr] * R.m[0][rr]) * dS[c][0] + 2.0f * (s[rr] * R.m[1][rr]) * dS[c][1] +
                             2.0f * (s[rr] * R.m[2][rr]) * dS[c][2];
        o.d_scale[0] = R.m[0][0] * dMt[0][0] + R.m[1][0] * dMt[0][1] + R.m[2][0] * dMt[0][2];
        o.d_scale[1] = R.m[0][1] * dMt[1][0] + R.m[1][1] * dMt[1][1] + R.m[2][1] * dMt[1][2];
        o.d_scale[2] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dMt[0][k] *= s[0];
            dMt[1][k] *= s[1];
            dMt[2][k] *= s[2];
        }
        const float gnx = g16[9], gny = g16[10], gnz = g16[11];
        dMt[2][0] += gnx * V[0] + gny * V[1] + gnz * V[2];
        dMt[2][1] += gnx * V[4] + gny * V[5] + gnz * V[6];
        dMt[2][2] += gnx * V[8] + gny * V[9] + gnz * V[10];
        o.d_rot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
        o.d_rot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                     4 * x * (dMt[2][2] + dMt[1][1]);
        o.d_rot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                     4 * y * (dMt[2][2] + dMt[0][0]);
        o.d_rot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                     4 * z * (dMt[1][1] + dMt[0][0]);
    }
}

// ---- conservative footprint bound (sharded projection, egs_preprocess.cu: k_surfel_candidates) -----------------------
// A tile rectangle that CONTAINS the exact one surfel_forward would compute (tests/test_hostemu_cpu.py checks that on
// the CPU against the exact rectangles of random and adversarial surfels), from ~70 instructions instead of ~1700:
//   * centre: the same projection as the exact path, up to rounding (2 px of slack below);
//   * radius: cov2D = T^T Sigma T + 0.3 I with T = W J, so lambda_max(cov2D) <= |J|_F^2 |R|^2 (mod max(|sx|, |sy|))^2 + 0.3,
//     where |J|_F^2 <= (fx/z)^2 (1 + lx^2) + (fy/z)^2 (1 + ly^2), lx / ly = 1.3 tanfov (the clamp of the reference's
//     computeCov2D, forward.cu:93-98), W the rigid part of the view matrix (norm 1, as the reference's own frustum test
//     assumes) and |R| <= |1 - |q|^2| + |q|^2 for the rotation built from a quaternion that is not normalised
//     (R(q) = (1 - |q|^2) I + |q|^2 R(q / |q|); 1 for unit quaternions).  The reference's radius
//     ceil(3 sqrt(lambda_1)), lambda_1 = mid + sqrt(max(0.1, mid^2 - det)), never exceeds 3 sqrt(lambda_max + 0.3163).
// Returns false when it declines to decide (behind / near the camera, absurd radius): the caller keeps the surfel.
EGS_HD bool surfel_bound_rect(const FrameConst& fc, const float* mean, const float* scale, const float* rot, int& x0,
                              int& y0, int& x1, int& y1) {
    const float px = mean[0], py = mean[1], pz = mean[2];
    const float hw = xf_affine(fc.proj, 3, px, py, pz);
    const float vz = xf_affine(fc.view, 2, px, py, pz);
    if (!(vz > 0.05f) || !(hw > 0.05f)) return false;
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ix = xf_affine(fc.proj, 0, px, py, pz) * pw * (float)fc.W * 0.5f + fc.cx;
    const float iy = xf_affine(fc.proj, 1, px, py, pz) * pw * (float)fc.H * 0.5f + fc.cy;
    const float q2 = rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3];
    const float rn = (fabsf(1.f - q2) + q2) * 1.0001f;
    const float smax = fc.mod * rn * fmaxf(fabsf(scale[0]), fabsf(scale[1]));
    const float lx = 1.3f * fc.tanfovx, ly = 1.3f * fc.tanfovy, iz = 1.0f / vz;
    const float jf = (fc.fx * iz) * (fc.fx * iz) * (1.f + lx * lx) + (fc.fy * iz) * (fc.fy * iz) * (1.f + ly * ly);
    const float lam = smax * smax * jf + 0.3f;
    const float rad = ceilf(3.f * sqrtf(lam + 0.32f) * 1.001f) + 2.f;   // + 2 px: the centre above is not the exact one
    if (!(rad < 16384.f)) return false;                                    // also catches NaN / Inf inputs
    egs_tile_rect(ix, iy, (int)rad, fc.gx, fc.gy, x0, y0, x1, y1);
    return true;
}
