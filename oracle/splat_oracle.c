/*
 * splat_oracle.c -- CPU restatement of the reference surfel rasterizer (forward + backward).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (eggfusion_b200/) may call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and only as the checker / CPU baseline.
 *
 * It restates, function by function, the algorithm of
 *   /root/reference/submodules/diff-gaussian-surfels/cuda_rasterizer/{forward,backward,rasterizer_impl}.cu
 *   /root/reference/submodules/diff-gaussian-surfels/cuda_rasterizer/auxiliary.h
 * (abbreviated DGS/... below) in plain C.  Parity pins: tests/golden/*.npz, produced on a B200 by the
 * unmodified reference compiled from source (oracle/build_ref.sh -> oracle/_ref, tests/golden/make_golden.py).
 *
 * Floating point: build with -ffp-contract=off.  The index-critical chain (everything that decides
 * radii / tile rectangles / cull tests / depth sort keys) spells out each rounding explicitly with
 * fmaf() so that it reproduces the FFMA grouping nvcc 12.9 emits for the reference on sm_100a
 * (read off its SASS: `a*b + c*d` -> fma(a,b,fl(c*d)); `x*y + z` / `z + x*y` -> fma(x,y,z);
 * `x*y - z` -> fma(x,y,-z); `z - x*y` -> fma(-x,y,z); IEEE-rounded div / sqrt / rcp).
 * Everything downstream of the index artefacts (colours, depth, gradients) is compared with a tolerance.
 *
 * Matrix convention (DGS/.../auxiliary.h:59-98): `view` and `proj` are the 16 floats of the transposed
 * (row-vector) matrices as EGG-Fusion passes them, i.e. element m[4*c + r] multiplies input component c
 * and contributes to output component r.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EGSO_TILE 16 /* DGS/cuda_rasterizer/config.h:15-17 */

typedef struct egso_camera {
    int32_t W, H;
    int32_t sh_degree;  /* active degree D */
    int32_t sh_coeffs;  /* M = coefficients stored per surfel; 0 => colours are given directly */
    float tanfovx, tanfovy;
    float cx, cy;
    float scale_modifier;
    float bg[3];
    float view[16];
    float proj[16];
    float campos[3];
} egso_camera;

/* ---------------------------------------------------------------- small helpers */

static const float SH0 = 0.28209479177387814f;
static const float SH1 = 0.4886025119029199f;
static const float SH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
static const float SH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* a*b + c*d + e*f as nvcc contracts it: second product rounded, first and third fused. */
static inline float dot3c(float a, float b, float c, float d, float e, float f) {
    float t = c * d;
    t = fmaf(a, b, t);
    return fmaf(e, f, t);
}

/* one output component of transformPoint4x3/4x4 (auxiliary.h:59-78) */
static inline float affine_row(const float* m, int r, float x, float y, float z) {
    return dot3c(m[r], x, m[4 + r], y, m[8 + r], z) + m[12 + r];
}

/* one output component of transformVec4x3 (auxiliary.h:80-88) */
static inline float linear_row(const float* m, int r, float x, float y, float z) {
    return dot3c(m[r], x, m[4 + r], y, m[8 + r], z);
}

static inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* getRect (auxiliary.h:47-57): float arithmetic, truncation toward zero, clamp to the tile grid. */
static void tile_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    float r = (float)radius;
    *x0 = imin(gx, imax(0, (int)((px - r) / (float)EGSO_TILE)));
    *y0 = imin(gy, imax(0, (int)((py - r) / (float)EGSO_TILE)));
    *x1 = imin(gx, imax(0, (int)((((px + r) + (float)EGSO_TILE) - 1.0f) / (float)EGSO_TILE)));
    *y1 = imin(gy, imax(0, (int)((((py + r) + (float)EGSO_TILE) - 1.0f) / (float)EGSO_TILE)));
}

/* quaternion2rotmat (forward.cu:115-129).  Returned as Rg[i][j] == glm's R[i][j] (column i, row j),
 * numerically the usual rotation matrix entry (row i, col j). */
static void quat_to_Rg(const float* q, float Rg[3][3]) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    Rg[0][0] = 1.f - 2.f * fmaf(y, y, z * z);
    Rg[0][1] = 2.f * fmaf(x, y, -(r * z));
    Rg[0][2] = 2.f * fmaf(x, z, r * y);
    Rg[1][0] = 2.f * fmaf(x, y, r * z);
    Rg[1][1] = 1.f - 2.f * fmaf(x, x, z * z);
    Rg[1][2] = 2.f * fmaf(y, z, -(r * x));
    Rg[2][0] = 2.f * fmaf(x, z, -(r * y));
    Rg[2][1] = 2.f * fmaf(y, z, r * x);
    Rg[2][2] = 1.f - 2.f * fmaf(x, x, y * y);
}

/* normalize() of auxiliary.h:197-203: returns the modulus, divides in place. */
static float normalize3(float* v) {
    float m = fmaxf(sqrtf(dot3c(v[0], v[0], v[1], v[1], v[2], v[2])), 0.00000001f);
    v[0] /= m;
    v[1] /= m;
    v[2] /= m;
    return m;
}

/* local_homo (auxiliary.h:205-281).  Returns 1 when the surfel is seen at a grazing angle (culled),
 * else fills J[0..3] (pixel -> tangent-plane 2x2 map) and J[4..6] = ax0, J[7..9] = ax1. */
static int local_homography(const float pv[3], const float nv[3], float fx, float fy, const float ax0[3],
                            const float ax1[3], float J[10]) {
    float prx = pv[0] / pv[2], pry = pv[1] / pv[2];
    const float s_fix = 1000.f;
    float svp = (fx + fy) / 2.f;
    float d0[3] = {prx + 1.f / s_fix, pry, 1.f};
    float m0 = normalize3(d0);
    float d1[3] = {prx, pry + 1.f / s_fix, 1.f};
    float m1 = normalize3(d1);
    float prj0 = dot3c(d0[0], nv[0], d0[1], nv[1], d0[2], nv[2]);
    float prj1 = dot3c(d1[0], nv[0], d1[1], nv[1], d1[2], nv[2]);
    if (fabsf(prj0 / m0) < 0.01f || fabsf(prj1 / m1) < 0.01f) return 1;

    float tt = dot3c(pv[0], nv[0], pv[1], nv[1], pv[2], nv[2]);
    float t0 = tt / prj0, t1 = tt / prj1;
    float xu0[3], xu1[3];
    for (int i = 0; i < 3; i++) {
        xu0[i] = fmaf(d0[i], t0, -pv[i]);
        xu1[i] = fmaf(d1[i], t1, -pv[i]);
    }
    /* the Surface-Splatting u0/u1 of auxiliary.h:243-249 is overwritten by ax0/ax1 (:252-257) */
    float k = svp / s_fix;
    J[0] = dot3c(xu0[0], ax0[0], xu0[1], ax0[1], xu0[2], ax0[2]) / k;
    J[1] = dot3c(xu1[0], ax0[0], xu1[1], ax0[1], xu1[2], ax0[2]) / k;
    J[2] = dot3c(xu0[0], ax1[0], xu0[1], ax1[1], xu0[2], ax1[2]) / k;
    J[3] = dot3c(xu1[0], ax1[0], xu1[1], ax1[1], xu1[2], ax1[2]) / k;
    for (int i = 0; i < 3; i++) {
        J[4 + i] = ax0[i];
        J[7 + i] = ax1[i];
    }
    return 0;
}

/* computeCov3D forward (forward.cu:135-155): S = diag(mod*sx, mod*sy, 0); Sigma = (S Rg)^T (S Rg). */
static void cov3d_from_scale_rot(const float* scale, float mod, float Rg[3][3], float* c6) {
    float M[3][2]; /* M[c][r], r = 0,1 (row 2 is exactly zero) */
    float s0 = mod * scale[0], s1 = mod * scale[1];
    for (int c = 0; c < 3; c++) {
        M[c][0] = s0 * Rg[c][0];
        M[c][1] = s1 * Rg[c][1];
    }
#define SIG(c, r) fmaf(M[r][0], M[c][0], M[r][1] * M[c][1])
    c6[0] = SIG(0, 0);
    c6[1] = SIG(0, 1);
    c6[2] = SIG(0, 2);
    c6[3] = SIG(1, 1);
    c6[4] = SIG(1, 2);
    c6[5] = SIG(2, 2);
#undef SIG
}

/* T = W * J of computeCov2D (forward.cu:74-99 and backward.cu:166-192); fills T[c][r] for c = 0,1
 * (column 2 is zero), and returns the clamped t.  txtz / tytz report the unclamped ratios. */
static void ewa_T(const float pv[3], float fx, float fy, float tanx, float tany, const float* view, float T[2][3],
                  float t[3], float* txtz, float* tytz) {
    float limx = 1.3f * tanx, limy = 1.3f * tany;
    *txtz = pv[0] / pv[2];
    *tytz = pv[1] / pv[2];
    t[0] = fminf(limx, fmaxf(-limx, *txtz)) * pv[2];
    t[1] = fminf(limy, fmaxf(-limy, *tytz)) * pv[2];
    t[2] = pv[2];
    float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
    float J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* W[c][r] = view[c + 4*r]  (forward.cu:94-97) */
    for (int r = 0; r < 3; r++) {
        float W0 = view[0 + 4 * r], W1 = view[1 + 4 * r], W2 = view[2 + 4 * r];
        T[0][r] = fmaf(W2, J02, W0 * J00);
        T[1][r] = fmaf(W2, J12, W1 * J11);
    }
}

/* cov2D = T^T Vrk^T T (+0.3 on the diagonal), forward.cu:101-112. */
static void ewa_cov2d(float T[2][3], const float* c6, float* a, float* b, float* c) {
    const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float A[3][2]; /* A[c][r] = sum_k T[r][k] V[c][k] */
    for (int cc = 0; cc < 3; cc++)
        for (int r = 0; r < 2; r++) A[cc][r] = dot3c(T[r][0], V[cc][0], T[r][1], V[cc][1], T[r][2], V[cc][2]);
    /* cov[c][r] = sum_k A[k][r] T[c][k] */
    *a = dot3c(A[0][0], T[0][0], A[1][0], T[0][1], A[2][0], T[0][2]) + 0.3f;
    *b = dot3c(A[0][1], T[0][0], A[1][1], T[0][1], A[2][1], T[0][2]);
    *c = dot3c(A[0][1], T[1][0], A[1][1], T[1][1], A[2][1], T[1][2]) + 0.3f;
}

/* computeColorFromSH forward (forward.cu:20-71) */
static void sh_to_rgb(int deg, const float* sh /* [M][3] */, const float pos[3], const float cam[3], float rgb[3],
                      uint8_t clamped[3]) {
    float d[3] = {pos[0] - cam[0], pos[1] - cam[1], pos[2] - cam[2]};
    float len = sqrtf(dot3c(d[0], d[0], d[1], d[1], d[2], d[2]));
    float x = d[0] / len, y = d[1] / len, z = d[2] / len;
    for (int ch = 0; ch < 3; ch++) {
#define S(k) sh[3 * (k) + ch]
        float res = SH0 * S(0);
        if (deg > 0) {
            res = fmaf(-(SH1 * y), S(1), res);
            res = fmaf(SH1 * z, S(2), res);
            res = fmaf(-(SH1 * x), S(3), res);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = fmaf(SH2[0] * xy, S(4), res);
                res = fmaf(SH2[1] * yz, S(5), res);
                res = fmaf(SH2[2] * (fmaf(2.0f, zz, -xx) - yy), S(6), res);
                res = fmaf(SH2[3] * xz, S(7), res);
                res = fmaf(SH2[4] * (xx - yy), S(8), res);
                if (deg > 2) {
                    res = fmaf(SH3[0] * y * fmaf(3.0f, xx, -yy), S(9), res);
                    res = fmaf(SH3[1] * xy * z, S(10), res);
                    res = fmaf(SH3[2] * y * (fmaf(4.0f, zz, -xx) - yy), S(11), res);
                    res = fmaf(SH3[3] * z * fmaf(-3.0f, yy, fmaf(2.0f, zz, -(3.0f * xx))), S(12), res);
                    res = fmaf(SH3[4] * x * (fmaf(4.0f, zz, -xx) - yy), S(13), res);
                    res = fmaf(SH3[5] * z * (xx - yy), S(14), res);
                    res = fmaf(SH3[6] * x * fmaf(-3.0f, yy, xx), S(15), res);
                }
            }
        }
#undef S
        res += 0.5f;
        clamped[ch] = (res < 0.f);
        rgb[ch] = fmaxf(res, 0.f);
    }
}

/* ---------------------------------------------------------------- forward: per-surfel stage */

/*
 * FORWARD::preprocess -> preprocessCUDA (forward.cu:158-301).
 * All outputs are caller-allocated and must be zero-initialised (the reference zero-fills radii /
 * active_mask / tiles_touched and leaves the rest of its geometry buffer untouched for culled surfels).
 *   radii[P] i32, active[P] u8, xy[P][2], depth[P], cov3d[P][6], conic_opacity[P][4], rgb[P][3],
 *   normal[P][3], jinv[P][10], viewcos[P], clamped[P][3] u8, tiles_touched[P] u32
 */
void egso_preprocess(const egso_camera* cam, int P, const float* means, const float* scales, const float* rots,
                     const float* opac, const float* shs, const float* colors_precomp, const int32_t* tile_mask,
                     int32_t* radii, uint8_t* active, float* xy, float* depth, float* cov3d, float* conic_opacity,
                     float* rgb, float* normal, float* jinv, float* viewcos, uint8_t* clamped,
                     uint32_t* tiles_touched) {
    const int W = cam->W, H = cam->H;
    const int gx = (W + EGSO_TILE - 1) / EGSO_TILE, gy = (H + EGSO_TILE - 1) / EGSO_TILE;
    const float fy = (float)H / (2.0f * cam->tanfovy), fx = (float)W / (2.0f * cam->tanfovx); /* rasterizer_impl.cu:244-245 */
    const float* V = cam->view;
    const float* PM = cam->proj;

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        float px = means[3 * i], py = means[3 * i + 1], pz = means[3 * i + 2];
        float hx = affine_row(PM, 0, px, py, pz), hy = affine_row(PM, 1, px, py, pz);
        float hw = affine_row(PM, 3, px, py, pz);
        float pw = 1.0f / (hw + 0.0000001f);
        float ndx = hx * pw, ndy = hy * pw;
        float pv[3] = {affine_row(V, 0, px, py, pz), affine_row(V, 1, px, py, pz), affine_row(V, 2, px, py, pz)};

        /* ndc2Pix (auxiliary.h:42-45): fp32 product, then one fp64 fma, then back to fp32 */
        float ix = (float)fma((double)(ndx * (float)W), 0.5, (double)cam->cx);
        float iy = (float)fma((double)(ndy * (float)H), 0.5, (double)cam->cy);

        /* in_frustum (auxiliary.h:140-149) */
        {
            const float e = 0.05f;
            float x0 = (float)(-W) * e, x1 = (float)W * (1 + e), y0 = (float)(-H) * e, y1 = (float)H * (1 + e);
            if (pv[2] < 0 || ix < x0 || ix >= x1 || iy < y0 || iy >= y1) continue;
        }
        active[i] = 1;

        float Rg[3][3];
        quat_to_Rg(rots + 4 * i, Rg);
        float nv[3], a0[3], a1[3];
        for (int r = 0; r < 3; r++) {
            nv[r] = linear_row(V, r, Rg[0][2], Rg[1][2], Rg[2][2]);
            a0[r] = linear_row(V, r, Rg[0][0], Rg[1][0], Rg[2][0]);
            a1[r] = linear_row(V, r, Rg[0][1], Rg[1][1], Rg[2][1]);
        }
        /* front_facing (auxiliary.h:180-194) */
        float facing = dot3c(pv[0], nv[0], pv[1], nv[1], pv[2], nv[2]);
        if ((double)facing > -0.00001) continue;
        viewcos[i] = facing;
        normal[3 * i] = nv[0];
        normal[3 * i + 1] = nv[1];
        normal[3 * i + 2] = nv[2];

        float J[10];
        if (local_homography(pv, nv, fx, fy, a0, a1, J)) continue;
        memcpy(jinv + 10 * i, J, sizeof(J));

        float* c6 = cov3d + 6 * i;
        cov3d_from_scale_rot(scales + 3 * i, cam->scale_modifier, Rg, c6);

        float T[2][3], t[3], txtz, tytz, ca, cb, cc;
        ewa_T(pv, fx, fy, cam->tanfovx, cam->tanfovy, V, T, t, &txtz, &tytz);
        ewa_cov2d(T, c6, &ca, &cb, &cc);

        float det = fmaf(ca, cc, -(cb * cb));
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conx = cc * det_inv, cony = -cb * det_inv, conz = ca * det_inv;
        float mid = 0.5f * (ca + cc);
        float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float lam = fmaxf(mid + sq, mid - sq);
        int rad = (int)ceilf(3.f * sqrtf(lam));

        int x0, y0, x1, y1;
        tile_rect(ix, iy, rad, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;

        if (colors_precomp == NULL) {
            sh_to_rgb(cam->sh_degree, shs + (size_t)3 * cam->sh_coeffs * i, means + 3 * i, cam->campos, rgb + 3 * i,
                      clamped + 3 * i);
        } else {
            rgb[3 * i] = colors_precomp[3 * i];
            rgb[3 * i + 1] = colors_precomp[3 * i + 1];
            rgb[3 * i + 2] = colors_precomp[3 * i + 2];
        }
        depth[i] = pv[2];
        radii[i] = rad;
        xy[2 * i] = ix;
        xy[2 * i + 1] = iy;
        conic_opacity[4 * i] = conx;
        conic_opacity[4 * i + 1] = cony;
        conic_opacity[4 * i + 2] = conz;
        conic_opacity[4 * i + 3] = opac[i];
        uint32_t cnt = 0;
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) cnt += tile_mask[y * gx + x] != 0;
        tiles_touched[i] = cnt;
    }
}

/* ---------------------------------------------------------------- forward: binning */

typedef struct {
    uint64_t key;
    uint32_t val;
} kv_t;

/* LSD radix sort on the low `bits` bits, 8 bits per pass: stable, exactly what
 * cub::DeviceRadixSort::SortPairs(..., 0, 32 + bit) guarantees (rasterizer_impl.cu:331-338). */
static void radix_sort_kv(kv_t* a, kv_t* tmp, size_t n, int bits) {
    for (int shift = 0; shift < bits; shift += 8) {
        size_t hist[257] = {0};
        int nb = bits - shift < 8 ? bits - shift : 8;
        uint64_t mask = ((uint64_t)1 << nb) - 1;
        for (size_t i = 0; i < n; i++) hist[((a[i].key >> shift) & mask) + 1]++;
        for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; i++) tmp[hist[(a[i].key >> shift) & mask]++] = a[i];
        memcpy(a, tmp, n * sizeof(kv_t));
    }
}

/* getHigherMsb (rasterizer_impl.cu:35-50) */
static uint32_t higher_msb(uint32_t n) {
    uint32_t msb = 16, step = 16;
    while (step > 1) {
        step /= 2;
        if (n >> msb)
            msb += step;
        else
            msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* Number of (surfel, tile) instances = last element of the inclusive scan (rasterizer_impl.cu:307-311). */
int64_t egso_count_instances(int P, const uint32_t* tiles_touched) {
    int64_t n = 0;
    for (int i = 0; i < P; i++) n += tiles_touched[i];
    return n;
}

/*
 * duplicateWithKeys + SortPairs + identifyTileRanges + host compaction (rasterizer_impl.cu:70-142,307-366).
 *   keys_sorted[I] u64, point_list[I] u32, ranges[tiles][2] u32 (zero for empty tiles),
 *   tile_indices[tiles] i32 (ascending ids of non-empty tiles, rest -1).  Returns tile_num.
 */
int egso_bin(const egso_camera* cam, int P, const int32_t* radii, const float* xy, const float* depth,
             const uint32_t* tiles_touched, const int32_t* tile_mask, int64_t I, uint64_t* keys_sorted,
             uint32_t* point_list, uint32_t* ranges, int32_t* tile_indices) {
    const int gx = (cam->W + EGSO_TILE - 1) / EGSO_TILE, gy = (cam->H + EGSO_TILE - 1) / EGSO_TILE;
    const int tiles = gx * gy;
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(I ? I : 1));
    kv_t* tmp = (kv_t*)malloc(sizeof(kv_t) * (size_t)(I ? I : 1));
    size_t off = 0;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        tile_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                uint64_t t = (uint64_t)(y * gx + x);
                if (!tile_mask[t]) continue;
                kv[off].key = (t << 32) | f2u(depth[i]);
                kv[off].val = (uint32_t)i;
                off++;
            }
    }
    (void)tiles_touched;
    radix_sort_kv(kv, tmp, off, 32 + (int)higher_msb((uint32_t)tiles));
    for (size_t k = 0; k < off; k++) {
        keys_sorted[k] = kv[k].key;
        point_list[k] = kv[k].val;
    }
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles);
    for (size_t k = 0; k < off; k++) {
        uint32_t cur = (uint32_t)(kv[k].key >> 32);
        if (k == 0)
            ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(kv[k - 1].key >> 32);
            if (cur != prev) {
                ranges[2 * prev + 1] = (uint32_t)k;
                ranges[2 * cur] = (uint32_t)k;
            }
        }
        if (k == off - 1) ranges[2 * cur + 1] = (uint32_t)off;
    }
    int n = 0;
    for (int t = 0; t < tiles; t++) tile_indices[t] = -1;
    for (int t = 0; t < tiles; t++)
        if (ranges[2 * t] != ranges[2 * t + 1]) tile_indices[n++] = t;
    free(kv);
    free(tmp);
    return n;
}

/* ---------------------------------------------------------------- forward: compositing */

/* pos_dif.z of depth_differencing (auxiliary.h:283-290) */
static inline float plane_depth_offset(float dx, float dy, const float* J) {
    float u0 = fmaf(dx, J[0], dy * J[1]);
    float u1 = fmaf(dx, J[2], dy * J[3]);
    return fmaf(u0, J[6], u1 * J[9]);
}

/* power of the conic at pixel offset d (forward.cu:421-425, backward.cu:550-554) */
static inline float conic_power(float cx_, float cy_, float cz_, float dx, float dy) {
    float q = fmaf(cx_ * dx, dx, (cz_ * dy) * dy);
    float dist = fmaf((2.f * cy_) * dx, dy, q);
    return -0.5f * dist;
}

/*
 * FORWARD::render -> renderCUDA (forward.cu:306-497).  Output images must be zero-initialised
 * (rasterize_points.cu:72-75): pixels of tiles with an empty list are never written.
 *   out_color[3][H][W], out_normal[3][H][W], out_depth[H][W], out_opac[H][W],
 *   final_T[H*W], final_D[H*W], n_contrib[H*W] u32
 */
void egso_render_forward(const egso_camera* cam, int tile_num, const int32_t* tile_indices, const uint32_t* ranges,
                         const uint32_t* point_list, const float* xy, const float* rgb, const float* normal,
                         const float* depth, const float* conic_opacity, const float* jinv, float* out_color,
                         float* out_normal, float* out_depth, float* out_opac, float* final_T, float* final_D,
                         uint32_t* n_contrib) {
    const int W = cam->W, H = cam->H;
    const int gx = (W + EGSO_TILE - 1) / EGSO_TILE;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < tile_num; b++) {
        int tile = tile_indices[b];
        int tx = tile % gx, ty = tile / gx;
        uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < EGSO_TILE; ly++)
            for (int lx = 0; lx < EGSO_TILE; lx++) {
                int x = tx * EGSO_TILE + lx, y = ty * EGSO_TILE + ly;
                if (x >= W || y >= H) continue;
                size_t pix = (size_t)W * y + x;
                float pxf = (float)x, pyf = (float)y;
                float T = 1.0f, C[3] = {0, 0, 0}, N[3] = {0, 0, 0}, D = 0;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = r0; k < r1; k++) {
                    uint32_t id = point_list[k];
                    contributor++;
                    float dx = xy[2 * id] - pxf, dy = xy[2 * id + 1] - pyf;
                    const float* co = conic_opacity + 4 * id;
                    float power = conic_power(co[0], co[1], co[2], dx, dy);
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break; /* `done`: this candidate is NOT blended */
                    float w = alpha * T;
                    float dj = depth[id] - plane_depth_offset(dx, dy, jinv + 10 * id);
                    D = fmaf(dj, w, D);
                    for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(rgb[3 * id + ch], w, C[ch]);
                    for (int ch = 0; ch < 3; ch++) N[ch] = fmaf(normal[3 * id + ch], w, N[ch]);
                    T = test_T;
                    last = contributor;
                }
                T = fminf((float)(1 - 0.000001), T);
                final_T[pix] = T;
                n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = fmaf(T, cam->bg[ch], C[ch]);
                for (int ch = 0; ch < 3; ch++) out_normal[ch * HW + pix] = N[ch];
                out_depth[pix] = D / (1 - T);
                out_opac[pix] = 1 - T;
                final_D[pix] = D;
            }
    }
}

/* ---------------------------------------------------------------- backward: compositing */

static inline void atomic_add_d(double* p, double v) {
#pragma omp atomic
    *p += v;
}

/*
 * BACKWARD::render -> renderCUDA (backward.cu:419-676).  Per-surfel sums are accumulated in fp64
 * (the reference uses fp32 atomics in nondeterministic order) and rounded to fp32 once at the end.
 * Outputs (zero-initialised by this function):
 *   d_mean2d[P][3] (z unused), d_conic[P][4] (x,y,w used), d_opacity[P], d_color[P][3], d_normal[P][3], d_depth[P]
 */
void egso_render_backward(const egso_camera* cam, int P, int tile_num, const int32_t* tile_indices,
                          const uint32_t* ranges, const uint32_t* point_list, const float* xy, const float* rgb,
                          const float* normal, const float* depth, const float* conic_opacity, const float* jinv,
                          const float* final_T, const float* final_D, const uint32_t* n_contrib,
                          const float* dL_dcolor, const float* dL_dnormal, const float* dL_ddepth_px,
                          const float* dL_dopac_px, float* d_mean2d, float* d_conic, float* d_opacity, float* d_color,
                          float* d_normal, float* d_depth) {
    const int W = cam->W, H = cam->H;
    const int gx = (W + EGSO_TILE - 1) / EGSO_TILE;
    const size_t HW = (size_t)H * W;
    /* 13 accumulators per surfel: mean2d x,y | conic x,y,w | opacity | color 3 | normal 3 | depth */
    double* acc = (double*)calloc((size_t)(P ? P : 1) * 13, sizeof(double));
    const float ddelx = 0.5f * (float)W, ddely = 0.5f * (float)H;

#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < tile_num; b++) {
        int tile = tile_indices[b];
        int tx = tile % gx, ty = tile / gx;
        uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < EGSO_TILE; ly++)
            for (int lx = 0; lx < EGSO_TILE; lx++) {
                int x = tx * EGSO_TILE + lx, y = ty * EGSO_TILE + ly;
                if (x >= W || y >= H) continue;
                size_t pix = (size_t)W * y + x;
                float pxf = (float)x, pyf = (float)y;
                const float T_final = final_T[pix], D_final = final_D[pix];
                float T = T_final;
                const int last_contributor = (int)n_contrib[pix];
                float gC[3], gN[3], gD, gO;
                for (int ch = 0; ch < 3; ch++) gC[ch] = dL_dcolor[ch * HW + pix];
                for (int ch = 0; ch < 3; ch++) gN[ch] = dL_dnormal[ch * HW + pix];
                gD = dL_ddepth_px[pix];
                gO = dL_dopac_px[pix];
                float acc_c[3] = {0, 0, 0}, acc_n[3] = {0, 0, 0}, acc_d = 0;
                float last_alpha = 0, last_c[3] = {0, 0, 0}, last_n[3] = {0, 0, 0}, last_d = 0;
                float bg_dot = 0;
                for (int ch = 0; ch < 3; ch++) bg_dot += cam->bg[ch] * gC[ch];

                /* walk the list back to front; entry k has contributor index k - r0 */
                for (int64_t k = (int64_t)r1 - 1; k >= (int64_t)r0; k--) {
                    int contributor = (int)(k - r0);
                    if (contributor >= last_contributor) continue;
                    uint32_t id = point_list[k];
                    float dx = xy[2 * id] - pxf, dy = xy[2 * id + 1] - pyf;
                    const float* co = conic_opacity + 4 * id;
                    float power = conic_power(co[0], co[1], co[2], dx, dy);
                    if (power > 0.0f) continue;
                    float G = expf(power);
                    float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;

                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    const float* J = jinv + 10 * id;
                    double* a = acc + (size_t)13 * id;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        float c = rgb[3 * id + ch];
                        acc_c[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * acc_c[ch];
                        last_c[ch] = c;
                        dL_dalpha += (c - acc_c[ch]) * gC[ch];
                        atomic_add_d(a + 6 + ch, (double)(w * gC[ch]));
                    }
                    for (int ch = 0; ch < 3; ch++) {
                        float n = normal[3 * id + ch];
                        acc_n[ch] = last_alpha * last_n[ch] + (1.f - last_alpha) * acc_n[ch];
                        last_n[ch] = n;
                        dL_dalpha += (n - acc_n[ch]) * gN[ch];
                        atomic_add_d(a + 9 + ch, (double)(w * gN[ch] * 10)); /* backward.cu:604: x10 */
                    }
                    {
                        float d_cur = depth[id] - plane_depth_offset(dx, dy, J);
                        acc_d = last_alpha * last_d + (1.f - last_alpha) * acc_d;
                        last_d = d_cur;
                        float gDn = gD / (1.f - T_final);
                        float t = gD * D_final / (1.f - T_final) / (1.f - T_final) * -T_final / (1 - alpha) / T;
                        t += (d_cur - acc_d) * gDn;
                        atomic_add_d(a + 12, (double)(w * gDn));
                        dL_dalpha += t;
                    }
                    dL_dalpha *= T;
                    dL_dalpha += gO * T_final / (1 - alpha);
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;

                    float dL_ddist = dL_dalpha * co[3] * -0.5f * G;
                    float gx_ = dL_ddist * 2 * (co[0] * dx + co[1] * dy) * ddelx;
                    float gy_ = dL_ddist * 2 * (co[2] * dy + co[1] * dx) * ddely;
                    /* backward.cu:659-660: depth-differencing term, un-weighted, not scaled by W/2 */
                    gx_ += -gD * (J[6] * J[0] + J[9] * J[2]);
                    gy_ += -gD * (J[6] * J[1] + J[9] * J[3]);
                    atomic_add_d(a + 0, (double)gx_);
                    atomic_add_d(a + 1, (double)gy_);
                    atomic_add_d(a + 2, (double)(dL_ddist * (dx * dx)));
                    atomic_add_d(a + 3, (double)(dL_ddist * (dx * dy)));
                    atomic_add_d(a + 4, (double)(dL_ddist * (dy * dy)));
                    atomic_add_d(a + 5, (double)(G * dL_dalpha));
                }
            }
    }
    for (int i = 0; i < P; i++) {
        const double* a = acc + (size_t)13 * i;
        d_mean2d[3 * i] = (float)a[0];
        d_mean2d[3 * i + 1] = (float)a[1];
        d_mean2d[3 * i + 2] = 0.f;
        d_conic[4 * i] = (float)a[2];
        d_conic[4 * i + 1] = (float)a[3];
        d_conic[4 * i + 2] = 0.f;
        d_conic[4 * i + 3] = (float)a[4];
        d_opacity[i] = (float)a[5];
        for (int ch = 0; ch < 3; ch++) d_color[3 * i + ch] = (float)a[6 + ch];
        for (int ch = 0; ch < 3; ch++) d_normal[3 * i + ch] = (float)a[9 + ch];
        d_depth[i] = (float)a[12];
    }
    free(acc);
}

/* ---------------------------------------------------------------- backward: per-surfel stage */

/* dnormvdv (auxiliary.h:108-118) */
static void dnormalize3(const float v[3], const float dv[3], float out[3]) {
    float s2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float inv = 1.0f / sqrtf(s2 * s2 * s2);
    out[0] = ((s2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
    out[1] = (-v[0] * v[1] * dv[0] + (s2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
    out[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (s2 - v[2] * v[2]) * dv[2]) * inv;
}

/* computeColorFromSH backward (backward.cu:20-139): writes d_sh[M][3], adds the view-direction term to d_mean. */
static void sh_backward(int deg, int M, const float* sh, const float pos[3], const float cam[3],
                        const uint8_t clamped[3], const float d_rgb_in[3], float* d_sh, float d_mean[3]) {
    float dor[3] = {pos[0] - cam[0], pos[1] - cam[1], pos[2] - cam[2]};
    float len = sqrtf(dor[0] * dor[0] + dor[1] * dor[1] + dor[2] * dor[2]);
    float x = dor[0] / len, y = dor[1] / len, z = dor[2] / len;
    float g[3];
    for (int ch = 0; ch < 3; ch++) g[ch] = d_rgb_in[ch] * (clamped[ch] ? 0.f : 1.f);
    float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0}; /* dRGB/d(dir) per channel */
    (void)M;
#define S(k, ch) sh[3 * (k) + (ch)]
#define OUT(k, coef)                                                     \
    for (int ch = 0; ch < 3; ch++) d_sh[3 * (k) + ch] = (coef) * g[ch]
    OUT(0, SH0);
    if (deg > 0) {
        OUT(1, -SH1 * y);
        OUT(2, SH1 * z);
        OUT(3, -SH1 * x);
        for (int ch = 0; ch < 3; ch++) {
            dx[ch] = -SH1 * S(3, ch);
            dy[ch] = -SH1 * S(1, ch);
            dz[ch] = SH1 * S(2, ch);
        }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            OUT(4, SH2[0] * xy);
            OUT(5, SH2[1] * yz);
            OUT(6, SH2[2] * (2.f * zz - xx - yy));
            OUT(7, SH2[3] * xz);
            OUT(8, SH2[4] * (xx - yy));
            for (int ch = 0; ch < 3; ch++) {
                dx[ch] += SH2[0] * y * S(4, ch) + SH2[2] * 2.f * -x * S(6, ch) + SH2[3] * z * S(7, ch) +
                          SH2[4] * 2.f * x * S(8, ch);
                dy[ch] += SH2[0] * x * S(4, ch) + SH2[1] * z * S(5, ch) + SH2[2] * 2.f * -y * S(6, ch) +
                          SH2[4] * 2.f * -y * S(8, ch);
                dz[ch] += SH2[1] * y * S(5, ch) + SH2[2] * 2.f * 2.f * z * S(6, ch) + SH2[3] * x * S(7, ch);
            }
            if (deg > 2) {
                OUT(9, SH3[0] * y * (3.f * xx - yy));
                OUT(10, SH3[1] * xy * z);
                OUT(11, SH3[2] * y * (4.f * zz - xx - yy));
                OUT(12, SH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                OUT(13, SH3[4] * x * (4.f * zz - xx - yy));
                OUT(14, SH3[5] * z * (xx - yy));
                OUT(15, SH3[6] * x * (xx - 3.f * yy));
                for (int ch = 0; ch < 3; ch++) {
                    dx[ch] += (SH3[0] * S(9, ch) * 3.f * 2.f * xy + SH3[1] * S(10, ch) * yz +
                               SH3[2] * S(11, ch) * -2.f * xy + SH3[3] * S(12, ch) * -3.f * 2.f * xz +
                               SH3[4] * S(13, ch) * (-3.f * xx + 4.f * zz - yy) + SH3[5] * S(14, ch) * 2.f * xz +
                               SH3[6] * S(15, ch) * 3.f * (xx - yy));
                    dy[ch] += (SH3[0] * S(9, ch) * 3.f * (xx - yy) + SH3[1] * S(10, ch) * xz +
                               SH3[2] * S(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                               SH3[3] * S(12, ch) * -3.f * 2.f * yz + SH3[4] * S(13, ch) * -2.f * xy +
                               SH3[5] * S(14, ch) * -2.f * yz + SH3[6] * S(15, ch) * -3.f * 2.f * xy);
                    dz[ch] += (SH3[1] * S(10, ch) * xy + SH3[2] * S(11, ch) * 4.f * 2.f * yz +
                               SH3[3] * S(12, ch) * 3.f * (2.f * zz - xx - yy) +
                               SH3[4] * S(13, ch) * 4.f * 2.f * xz + SH3[5] * S(14, ch) * (xx - yy));
                }
            }
        }
    }
#undef S
#undef OUT
    float ddir[3] = {dx[0] * g[0] + dx[1] * g[1] + dx[2] * g[2], dy[0] * g[0] + dy[1] * g[1] + dy[2] * g[2],
                     dz[0] * g[0] + dz[1] * g[1] + dz[2] * g[2]};
    float dm[3];
    dnormalize3(dor, ddir, dm);
    d_mean[0] += dm[0];
    d_mean[1] += dm[1];
    d_mean[2] += dm[2];
}

/*
 * BACKWARD::preprocess = computeCov2DCUDA (backward.cu:144-274) then preprocessCUDA (backward.cu:358-416)
 * incl. computeColorFromSH bwd (:20-139) and computeCov3D bwd (:278-353).
 * Inputs: the screen-space gradients produced by egso_render_backward, the forward's cov3d / clamped / radii.
 * Outputs (zero-initialised here): d_means[P][3], d_cov3d[P][6], d_sh[P][M][3], d_scales[P][3], d_rots[P][4].
 * Only surfels with radii > 0 are touched.  Reproduces the reference's deviations:
 * S.z = mod*scale.z in the cov3D backward, no quaternion-normalisation Jacobian, normal-gradient injection.
 */
void egso_preprocess_backward(const egso_camera* cam, int P, const float* means, const float* scales,
                              const float* rots, const float* shs, const int32_t* radii, const float* cov3d,
                              const uint8_t* clamped, const float* d_mean2d, const float* d_conic,
                              const float* d_color, const float* d_normal, const float* d_depth, float* d_means,
                              float* d_cov3d, float* d_sh, float* d_scales, float* d_rots) {
    const int M = cam->sh_coeffs;
    const float fy = (float)cam->H / (2.0f * cam->tanfovy), fx = (float)cam->W / (2.0f * cam->tanfovx);
    const float* V = cam->view;
    const float* PM = cam->proj;
    memset(d_means, 0, sizeof(float) * 3 * (size_t)P);
    memset(d_cov3d, 0, sizeof(float) * 6 * (size_t)P);
    if (M > 0) memset(d_sh, 0, sizeof(float) * 3 * (size_t)M * P);
    memset(d_scales, 0, sizeof(float) * 3 * (size_t)P);
    memset(d_rots, 0, sizeof(float) * 4 * (size_t)P);

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        const float* mean = means + 3 * i;
        const float* c6 = cov3d + 6 * i;
        float gm[3]; /* running dL/dmean */

        /* ---- computeCov2DCUDA */
        {
            float gcon[3] = {d_conic[4 * i], d_conic[4 * i + 1], d_conic[4 * i + 3]};
            float pv[3] = {affine_row(V, 0, mean[0], mean[1], mean[2]), affine_row(V, 1, mean[0], mean[1], mean[2]),
                           affine_row(V, 2, mean[0], mean[1], mean[2])};
            float T[2][3], t[3], txtz, tytz;
            ewa_T(pv, fx, fy, cam->tanfovx, cam->tanfovy, V, T, t, &txtz, &tytz);
            const float limx = 1.3f * cam->tanfovx, limy = 1.3f * cam->tanfovy;
            const float xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
            const float ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
            float a, b, c;
            ewa_cov2d(T, c6, &a, &b, &c);
            const float Vr[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
            float denom = a * c - b * b;
            float da = 0, db = 0, dc = 0;
            float d2inv = 1.0f / ((denom * denom) + 0.0000001f);
            float* gc = d_cov3d + 6 * i;
            if (d2inv != 0) {
                da = d2inv * (-c * c * gcon[0] + 2 * b * c * gcon[1] + (denom - a * c) * gcon[2]);
                dc = d2inv * (-a * a * gcon[2] + 2 * a * b * gcon[1] + (denom - a * c) * gcon[0]);
                db = d2inv * 2 * (b * c * gcon[0] - (denom + 2 * b * b) * gcon[1] + a * b * gcon[2]);
                gc[0] = (T[0][0] * T[0][0] * da + T[0][0] * T[1][0] * db + T[1][0] * T[1][0] * dc);
                gc[3] = (T[0][1] * T[0][1] * da + T[0][1] * T[1][1] * db + T[1][1] * T[1][1] * dc);
                gc[5] = (T[0][2] * T[0][2] * da + T[0][2] * T[1][2] * db + T[1][2] * T[1][2] * dc);
                gc[1] = 2 * T[0][0] * T[0][1] * da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * db +
                        2 * T[1][0] * T[1][1] * dc;
                gc[2] = 2 * T[0][0] * T[0][2] * da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * db +
                        2 * T[1][0] * T[1][2] * dc;
                gc[4] = 2 * T[0][2] * T[0][1] * da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * db +
                        2 * T[1][1] * T[1][2] * dc;
            }
            float dT[2][3];
            for (int k = 0; k < 3; k++) {
                float tv0 = T[0][0] * Vr[k][0] + T[0][1] * Vr[k][1] + T[0][2] * Vr[k][2];
                float tv1 = T[1][0] * Vr[k][0] + T[1][1] * Vr[k][1] + T[1][2] * Vr[k][2];
                dT[0][k] = 2 * tv0 * da + tv1 * db;
                dT[1][k] = 2 * tv1 * dc + tv0 * db;
            }
            /* W[c][r] = view[c + 4 r];  dJ(c,r) = sum_k W[r][k] dT[c][k] */
            float dJ00 = V[0] * dT[0][0] + V[4] * dT[0][1] + V[8] * dT[0][2];
            float dJ02 = V[2] * dT[0][0] + V[6] * dT[0][1] + V[10] * dT[0][2];
            float dJ11 = V[1] * dT[1][0] + V[5] * dT[1][1] + V[9] * dT[1][2];
            float dJ12 = V[2] * dT[1][0] + V[6] * dT[1][1] + V[10] * dT[1][2];
            float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
            float dtx = xmul * -fx * tz2 * dJ02;
            float dty = ymul * -fy * tz2 * dJ12;
            float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
            /* transformVec4x3Transpose (auxiliary.h:90-98) */
            gm[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
            gm[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
            gm[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
        }

        /* ---- preprocessCUDA (bwd): screen-space mean and depth paths */
        {
            float hw = affine_row(PM, 3, mean[0], mean[1], mean[2]);
            float mw = 1.0f / (hw + 0.0000001f);
            float mul1 = (PM[0] * mean[0] + PM[4] * mean[1] + PM[8] * mean[2] + PM[12]) * mw * mw;
            float mul2 = (PM[1] * mean[0] + PM[5] * mean[1] + PM[9] * mean[2] + PM[13]) * mw * mw;
            float g2x = d_mean2d[3 * i], g2y = d_mean2d[3 * i + 1];
            float dmx = (PM[0] * mw - PM[3] * mul1) * g2x + (PM[1] * mw - PM[3] * mul2) * g2y;
            float dmy = (PM[4] * mw - PM[7] * mul1) * g2x + (PM[5] * mw - PM[7] * mul2) * g2y;
            float dmz = (PM[8] * mw - PM[11] * mul1) * g2x + (PM[9] * mw - PM[11] * mul2) * g2y;
            float gd = d_depth[i];
            gm[0] += dmx + gd * V[2];
            gm[1] += dmy + gd * V[6];
            gm[2] += dmz + gd * V[10];
        }

        /* ---- SH */
        if (shs != NULL && M > 0)
            sh_backward(cam->sh_degree, M, shs + (size_t)3 * M * i, mean, cam->campos, clamped + 3 * i, d_color + 3 * i,
                        d_sh + (size_t)3 * M * i, gm);
        d_means[3 * i] = gm[0];
        d_means[3 * i + 1] = gm[1];
        d_means[3 * i + 2] = gm[2];

        /* ---- computeCov3D (bwd) */
        if (scales != NULL) {
            const float* q = rots + 4 * i;
            float r = q[0], x = q[1], y = q[2], z = q[3];
            float Rg[3][3];
            quat_to_Rg(q, Rg);
            float s[3] = {cam->scale_modifier * scales[3 * i], cam->scale_modifier * scales[3 * i + 1],
                          cam->scale_modifier * scales[3 * i + 2]};
            /* M = S * Rg (glm): Mg[c][r] = s[r] * Rg[c][r] */
            float Mg[3][3];
            for (int c = 0; c < 3; c++)
                for (int rr = 0; rr < 3; rr++) Mg[c][rr] = s[rr] * Rg[c][rr];
            const float* g6 = d_cov3d + 6 * i;
            /* dL_dSigma (symmetric), glm column-major but symmetric so index order is irrelevant */
            float dS[3][3] = {{g6[0], 0.5f * g6[1], 0.5f * g6[2]},
                              {0.5f * g6[1], g6[3], 0.5f * g6[4]},
                              {0.5f * g6[2], 0.5f * g6[4], g6[5]}};
            /* dL_dM = (2 M) * dSigma : dM[c][r] = sum_k 2 Mg[k][r] dS[c][k] */
            float dM[3][3];
            for (int c = 0; c < 3; c++)
                for (int rr = 0; rr < 3; rr++)
                    dM[c][rr] = 2.0f * Mg[0][rr] * dS[c][0] + 2.0f * Mg[1][rr] * dS[c][1] + 2.0f * Mg[2][rr] * dS[c][2];
            /* Rt[i][j] = Rg[j][i]; dMt[i][j] = dM[j][i] */
            float dMt[3][3];
            for (int a = 0; a < 3; a++)
                for (int bb = 0; bb < 3; bb++) dMt[a][bb] = dM[bb][a];
            d_scales[3 * i] = Rg[0][0] * dMt[0][0] + Rg[1][0] * dMt[0][1] + Rg[2][0] * dMt[0][2];
            d_scales[3 * i + 1] = Rg[0][1] * dMt[1][0] + Rg[1][1] * dMt[1][1] + Rg[2][1] * dMt[1][2];
            d_scales[3 * i + 2] = 0;
            for (int k = 0; k < 3; k++) {
                dMt[0][k] *= s[0];
                dMt[1][k] *= s[1];
                dMt[2][k] *= s[2];
            }
            /* normal-gradient injection (backward.cu:333-341): W^T (3x3) applied to dL/dn_view */
            const float* gn = d_normal + 3 * i;
            dMt[2][0] += gn[0] * V[0] + gn[1] * V[1] + gn[2] * V[2];
            dMt[2][1] += gn[0] * V[4] + gn[1] * V[5] + gn[2] * V[6];
            dMt[2][2] += gn[0] * V[8] + gn[1] * V[9] + gn[2] * V[10];
            float* gq = d_rots + 4 * i;
            gq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            gq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                    4 * x * (dMt[2][2] + dMt[1][1]);
            gq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                    4 * y * (dMt[2][2] + dMt[0][0]);
            gq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                    4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* checkFrustum / in_frustum(idx, ...) used by markVisible (rasterizer_impl.cu:54-66, auxiliary.h:152-178) */
void egso_mark_visible(int P, const float* means, const float* view, const float* proj, uint8_t* present) {
    for (int i = 0; i < P; i++) {
        float x = means[3 * i], y = means[3 * i + 1], z = means[3 * i + 2];
        float hx = affine_row(proj, 0, x, y, z), hy = affine_row(proj, 1, x, y, z), hw = affine_row(proj, 3, x, y, z);
        float pw = 1.0f / (hw + 0.0000001f);
        float ndx = hx * pw, ndy = hy * pw;
        float vz = affine_row(view, 2, x, y, z);
        present[i] = !(vz <= 0.2f || (double)ndx < -1.3 || (double)ndx > 1.3 || (double)ndy < -1.3 || (double)ndy > 1.3);
    }
}

int egso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void egso_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ================================================================================================
 * Surfel fusion kernels: /root/reference/submodules/diff-gaussian-surfels/fuse_surfels.cu
 * ================================================================================================ */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int rn_int(float x) { return (int)nearbyintf(x); } /* __float2int_rn */
static inline int rd_int(float x) { return (int)floorf(x); }    /* __float2int_rd */

/* shared front part of both kernels: projection, frustum test, rotation, back-face test.
 * Returns 0 when the surfel is rejected; *stage = 1 once it passed the frustum test. */
static int fuse_common(const float* view, const float* proj, int wd, int ht, float cx, float cy, const float* p,
                       const float* q, float pc[3], float coord[2], float Rg[3][3], float zaxis[3]) {
    float hx = affine_row(proj, 0, p[0], p[1], p[2]), hy = affine_row(proj, 1, p[0], p[1], p[2]);
    float hw = affine_row(proj, 3, p[0], p[1], p[2]);
    float pw = 1.0f / (hw + 0.0000001f);
    for (int r = 0; r < 3; r++) pc[r] = affine_row(view, r, p[0], p[1], p[2]);
    coord[0] = (float)fma((double)((hx * pw) * (float)wd), 0.5, (double)cx);
    coord[1] = (float)fma((double)((hy * pw) * (float)ht), 0.5, (double)cy);
    const float e = 0.05f;
    float x0 = (float)(-wd) * e, x1 = (float)wd * (1 + e), y0 = (float)(-ht) * e, y1 = (float)ht * (1 + e);
    if (pc[2] < 0 || coord[0] < x0 || coord[0] >= x1 || coord[1] < y0 || coord[1] >= y1) return 0;
    quat_to_Rg(q, Rg);
    for (int r = 0; r < 3; r++) zaxis[r] = linear_row(view, r, Rg[0][2], Rg[1][2], Rg[2][2]);
    float facing = dot3c(pc[0], zaxis[0], pc[1], zaxis[1], pc[2], zaxis[2]);
    if ((double)facing > -0.00001) return 0;
    return 1;
}

/*
 * projectSurfelsToFrame (fuse_surfels.cu:475-536): z-buffer of stable, front-facing surfels splatted on a 3x3
 * pixel footprint.  The reference races (atomicMin on the depth, then a plain store of the index); this is the
 * race-free outcome: per pixel the smallest depth, ties to the smallest surfel index.
 * index_map[ht*wd] must be pre-filled with -1, depth_buffer[ht*wd] with +inf (reference __init__.py:316-317).
 */
void egso_project_surfels(int P, int ht, int wd, const float* points, const float* rotations, const uint8_t* stable,
                          const float* intrinsic, const float* view, const float* proj, int32_t* index_map,
                          float* depth_buffer) {
    const float cx = intrinsic[2], cy = intrinsic[3];
    for (int i = 0; i < P; i++) {
        if (!stable[i]) continue;
        float pc[3], coord[2], Rg[3][3], z[3];
        if (!fuse_common(view, proj, wd, ht, cx, cy, points + 3 * i, rotations + 4 * i, pc, coord, Rg, z)) continue;
        int x = (int)coord[0], y = (int)coord[1];
        if (x < 0 || x >= wd || y < 0 || y >= ht) continue;
        float depth = pc[2];
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int nx = x + dx, ny = y + dy;
                if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
                int k = ny * wd + nx;
                if (depth < depth_buffer[k]) {
                    depth_buffer[k] = depth;
                    index_map[k] = i;
                }
            }
    }
}

static int all4_nonzero_f(const float* map, int wd, int x, int y, int ch) {
    for (int c = 0; c < ch; c++) {
        if (map[(y * wd + x) * ch + c] == 0) return 0;
        if (map[(y * wd + x + 1) * ch + c] == 0) return 0;
        if (map[((y + 1) * wd + x) * ch + c] == 0) return 0;
        if (map[((y + 1) * wd + x + 1) * ch + c] == 0) return 0;
    }
    return 1;
}

/*
 * preprocessSurfel (fuse_surfels.cu:214-394): in-place information-filter fusion of surfel position / normal
 * with the current frame.  Mutates points[P][3], rotations[P][4] (raw, un-normalised quaternions), sigma2[P][2];
 * writes inview_mask[P], surface_mask[P].  frame_vmap / frame_nmap are [ht][wd][3], frame_dmap [ht][wd],
 * frame_mask [ht][wd] (bool bytes), frame_imap [ht][wd] i32.  The arguments the reference kernel receives but
 * never reads (scales, colors, confidence, tic, eta, counts, stable mask, depth buffer, model maps) are omitted.
 */
void egso_fuse_surfels(int P, int ht, int wd, const float* intrinsic, const float* view, const float* proj,
                       const float* frame_vmap, const float* frame_nmap, const float* frame_dmap,
                       const uint8_t* frame_mask, const int32_t* frame_imap, float* points, float* rotations,
                       float* sigma2, uint8_t* inview, uint8_t* surface, float dist_thres, float alpha_p,
                       float alpha_n) {
    const float cx = intrinsic[2], cy = intrinsic[3];
    for (int i = 0; i < P; i++) {
        inview[i] = 0;
        surface[i] = 0;
        float pw[3] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
        float pc[3], coord[2], R[3][3], zax[3];
        if (!fuse_common(view, proj, wd, ht, cx, cy, pw, rotations + 4 * i, pc, coord, R, zax)) continue;
        inview[i] = 1;
        {   /* is_valid on the normal map (3 channels) and on the mask (fuse_surfels.cu:166-212) */
            int x = rd_int(coord[0]), y = rd_int(coord[1]);
            if (x < 0 || x + 1 >= wd || y < 0 || y + 1 >= ht) continue;
            if (!all4_nonzero_f(frame_nmap, wd, x, y, 3)) continue;
            if (!frame_mask[y * wd + x] || !frame_mask[y * wd + x + 1] || !frame_mask[(y + 1) * wd + x] ||
                !frame_mask[(y + 1) * wd + x + 1])
                continue;
        }
        int ix = clampi(rn_int(coord[0]), 0, wd - 1), iy = clampi(rn_int(coord[1]), 0, ht - 1);
        int nidx = iy * wd + ix;
        float nw[3] = {R[0][2], R[1][2], R[2][2]};
        {
            float inv = 1.0f / sqrtf(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
            nw[0] *= inv; nw[1] *= inv; nw[2] *= inv;
        }
        float nc[3] = {frame_nmap[3 * nidx], frame_nmap[3 * nidx + 1], frame_nmap[3 * nidx + 2]};
        {
            float inv = 1.0f / sqrtf(nc[0] * nc[0] + nc[1] * nc[1] + nc[2] * nc[2]);
            nc[0] *= inv; nc[1] *= inv; nc[2] *= inv;
        }
        float vc[3] = {frame_vmap[3 * nidx], frame_vmap[3 * nidx + 1], frame_vmap[3 * nidx + 2]};
        float vd[3] = {pw[0] - vc[0], pw[1] - vc[1], pw[2] - vc[2]};
        int sid = frame_imap[rn_int(coord[1]) * wd + rn_int(coord[0])];
        if (sid >= 0 && sid == i) surface[i] = 1;
        if (sqrtf(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2]) > dist_thres) continue;

        float s2p = sigma2[2 * i], s2n = sigma2[2 * i + 1];
        float eta_p[3] = {pw[0] / s2p, pw[1] / s2p, pw[2] / s2p};
        float eta_n[3] = {nw[0] / s2n, nw[1] / s2n, nw[2] / s2n};
        float d = frame_dmap[nidx];
        float s2pz = (alpha_p * d) * (alpha_p * d), s2nz = (alpha_n * d) * (alpha_n * d);
        float s2p_new = 1.f / (1.f / s2pz + 1.f / s2p), s2n_new = 1.f / (1.f / s2nz + 1.f / s2n);
        float lp = 1.f / s2pz, ln = 1.f / s2nz;
        float xn[3];
        for (int k = 0; k < 3; k++) {
            points[3 * i + k] = s2p_new * (eta_p[k] + lp * vc[k]);
            xn[k] = s2n_new * (eta_n[k] + ln * nc[k]);
        }
        sigma2[2 * i] = s2p_new;

        float dotn = nw[0] * nc[0] + nw[1] * nc[1] + nw[2] * nc[2];
        dotn = dotn > 1 ? 1 : (dotn < -1 ? -1 : dotn);
        double angle = (double)(acosf(dotn) * 180) / 3.1415926;
        if (!(angle < 60)) continue;
        float inv = 1.0f / sqrtf(xn[0] * xn[0] + xn[1] * xn[1] + xn[2] * xn[2]);
        float nn[3] = {xn[0] * inv, xn[1] * inv, xn[2] * inv};
        float cr[3] = {nw[1] * nn[2] - nw[2] * nn[1], nw[2] * nn[0] - nw[0] * nn[2], nw[0] * nn[1] - nw[1] * nn[0]};
        inv = 1.0f / sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        float n12[3] = {cr[0] * inv, cr[1] * inv, cr[2] * inv};
        float ct = nw[0] * nn[0] + nw[1] * nn[1] + nw[2] * nn[2];
        ct = ct > 1 ? 1 : (ct < -1 ? -1 : ct);
        double theta = (double)(acosf(ct) * 180) / 3.1415926;
        if (theta < 1) continue;
        /* rotvec2rotmatrix (fuse_surfels.cu:64-89); the vector is normalised again inside */
        float ang = acosf(ct);
        float nrm = sqrtf(n12[0] * n12[0] + n12[1] * n12[1] + n12[2] * n12[2]);
        float ux = n12[0] / nrm, uy = n12[1] / nrm, uz = n12[2] / nrm;
        float c = cosf(ang), s = sinf(ang);
        float R1[3][3];
        R1[0][0] = c + ux * ux * (1 - c);      R1[0][1] = ux * uy * (1 - c) - uz * s; R1[0][2] = ux * uz * (1 - c) + uy * s;
        R1[1][0] = uy * ux * (1 - c) + uz * s; R1[1][1] = c + uy * uy * (1 - c);      R1[1][2] = uy * uz * (1 - c) - ux * s;
        R1[2][0] = uz * ux * (1 - c) - uy * s; R1[2][1] = uz * uy * (1 - c) + ux * s; R1[2][2] = c + uz * uz * (1 - c);
        /* glm R * R1 with both stored "transposed": in usual indexing R2 = R1 . R */
        float R2[3][3];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) R2[a][b] = R[0][b] * R1[a][0] + R[1][b] * R1[a][1] + R[2][b] * R1[a][2];
        /* rotmat2quaternion (fuse_surfels.cu:31-62) */
        float tr = R2[0][0] + R2[1][1] + R2[2][2], qw, qx, qy, qz;
        if (tr > 0.0f) {
            float ss = 0.5f / sqrtf(tr + 1.0f);
            qw = 0.25f / ss; qx = (R2[2][1] - R2[1][2]) * ss; qy = (R2[0][2] - R2[2][0]) * ss; qz = (R2[1][0] - R2[0][1]) * ss;
        } else if (R2[0][0] > R2[1][1] && R2[0][0] > R2[2][2]) {
            float ss = 2.0f * sqrtf(1.0f + R2[0][0] - R2[1][1] - R2[2][2]);
            qw = (R2[2][1] - R2[1][2]) / ss; qx = 0.25f * ss; qy = (R2[0][1] + R2[1][0]) / ss; qz = (R2[0][2] + R2[2][0]) / ss;
        } else if (R2[1][1] > R2[2][2]) {
            float ss = 2.0f * sqrtf(1.0f + R2[1][1] - R2[0][0] - R2[2][2]);
            qw = (R2[0][2] - R2[2][0]) / ss; qx = (R2[0][1] + R2[1][0]) / ss; qy = 0.25f * ss; qz = (R2[1][2] + R2[2][1]) / ss;
        } else {
            float ss = 2.0f * sqrtf(1.0f + R2[2][2] - R2[0][0] - R2[1][1]);
            qw = (R2[1][0] - R2[0][1]) / ss; qx = (R2[0][2] + R2[2][0]) / ss; qy = (R2[1][2] + R2[2][1]) / ss; qz = 0.25f * ss;
        }
        rotations[4 * i] = qw; rotations[4 * i + 1] = qx; rotations[4 * i + 2] = qy; rotations[4 * i + 3] = qz;
        sigma2[2 * i + 1] = s2n_new;
    }
}

/* ================================================================================================
 * Dense-tracking image utilities: /root/reference/src/utils/cuda/src/tracking.cu ("TRK")
 * ================================================================================================ */

/* bilateral_filter_kernel (TRK:777-821) */
void egto_bilateral(const float* in, float* out, int wd, int ht, int window, float sigma_c, float sigma_s) {
    const float ss = 1.0f / (2.0f * sigma_s * sigma_s), sc = 1.0f / (2.0f * sigma_c * sigma_c);
    const int r = window / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ht; y++)
        for (int x = 0; x < wd; x++) {
            float c = in[y * wd + x], s1 = 0.f, s2 = 0.f;
            for (int dy = -r; dy <= r; dy++)
                for (int dx = -r; dx <= r; dx++) {
                    int nx = x + dx, ny = y + dy;
                    if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
                    float v = in[ny * wd + nx], dc = c - v;
                    float space2 = (float)(dx * dx + dy * dy), color2 = dc * dc;
                    float w = expf(-space2 * ss - color2 * sc);
                    s1 += v * w;
                    s2 += w;
                }
            out[y * wd + x] = s1 / s2;
        }
}

/* gaussian_filter_kernel (TRK:705-751) */
void egto_gaussian(const float* in, float* out, int wd, int ht, int ch, int window, float sigma_s) {
    const float ss = 1.0f / (2.0f * sigma_s * sigma_s);
    const int r = window / 2;
    for (int y = 0; y < ht; y++)
        for (int x = 0; x < wd; x++) {
            float s1[4] = {0, 0, 0, 0}, s2 = 0.f;
            for (int dy = -r; dy <= r; dy++)
                for (int dx = -r; dx <= r; dx++) {
                    int nx = x + dx, ny = y + dy;
                    if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
                    float w = expf(-(float)(dx * dx + dy * dy) * ss);
                    for (int c = 0; c < ch; c++) s1[c] += in[(ny * wd + nx) * ch + c] * w;
                    s2 += w;
                }
            for (int c = 0; c < ch; c++) out[(y * wd + x) * ch + c] = s1[c] / s2;
        }
}

/* gaussian_downsample_kernel (TRK:533-575) with the 5x5 table of TRK:585-586 */
void egto_downsample(const float* in, float* out, int wd, int ht, int ch) {
    static const float k[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};
    const int dw = wd / 2, dh = ht / 2;
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            float sum[4] = {0, 0, 0, 0}, count = 0.f;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    int nx = 2 * x + dx, ny = 2 * y + dy;
                    if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
                    float w = k[(dy + 2) * 5 + (dx + 2)];
                    for (int c = 0; c < ch; c++) sum[c] += in[(ny * wd + nx) * ch + c] * w;
                    count += w;
                }
            for (int c = 0; c < ch; c++) out[(y * dw + x) * ch + c] = sum[c] / count;
        }
}

/* gradient_kernel (TRK:853-893) with the tables of TRK:903-909, walked from index 8 down to 0 */
void egto_gradients(const float* in, float* gx, float* gy, int wd, int ht) {
    static const float kx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    static const float ky[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
    for (int y = 0; y < ht; y++)
        for (int x = 0; x < wd; x++) {
            float ax = 0.f, ay = 0.f;
            int k = 8;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int nx = x + dx, ny = y + dy;
                    if (nx >= 0 && nx < wd && ny >= 0 && ny < ht) {
                        float v = in[ny * wd + nx];
                        ax += v * kx[k];
                        ay += v * ky[k];
                    }
                    --k;
                }
            gx[y * wd + x] = ax;
            gy[y * wd + x] = ay;
        }
}

/* compute_vertex_map_kernel + compute_normal_map_kernel (TRK:602-672) */
void egto_vertex_normal(const float* depth, float fx, float fy, float cx, float cy, float* vmap, float* nmap, int wd,
                        int ht) {
    for (int y = 0; y < ht; y++)
        for (int x = 0; x < wd; x++) {
            int i = y * wd + x;
            float Z = depth[i];
            vmap[3 * i] = ((float)x - cx) * Z / fx;
            vmap[3 * i + 1] = ((float)y - cy) * Z / fy;
            vmap[3 * i + 2] = Z;
        }
    for (int y = 0; y < ht; y++)
        for (int x = 0; x < wd; x++) {
            int i = y * wd + x;
            const float* v00 = vmap + 3 * i;
            const float* v10 = (x + 1 < wd) ? vmap + 3 * (i + 1) : v00;
            const float* v01 = (y + 1 < ht) ? vmap + 3 * (i + wd) : v00;
            float a[3] = {v01[0] - v00[0], v01[1] - v00[1], v01[2] - v00[2]};
            float b[3] = {v10[0] - v00[0], v10[1] - v00[1], v10[2] - v00[2]};
            /* nvcc contracts x*y - u*v into fma(x, y, -(u*v)): for parallel a, b (both neighbours are holes) the
             * result is the rounding error of one product, not 0, and the reference then emits a unit vector */
            float n[3] = {fmaf(a[1], b[2], -(a[2] * b[1])), fmaf(a[2], b[0], -(a[0] * b[2])),
                          fmaf(a[0], b[1], -(a[1] * b[0]))};
            float d2 = fmaf(n[2], n[2], fmaf(n[0], n[0], n[1] * n[1]));
            if (d2 < 1.17549435e-38f) d2 = 0.f; /* rsqrtf (rsqrt.approx.ftz) flushes denormal inputs to zero */
            float inv = 1.0f / sqrtf(d2);
            n[0] *= inv; n[1] *= inv; n[2] *= inv;
            if (isnan(n[0]) || isnan(n[1]) || isnan(n[2])) n[0] = n[1] = n[2] = 0.f;
            nmap[3 * i] = n[0]; nmap[3 * i + 1] = n[1]; nmap[3 * i + 2] = n[2];
        }
}
