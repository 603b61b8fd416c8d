// hostemu.cpp -- TEST ONLY.  Compiles the product's __host__ __device__ per-surfel math
// (eggfusion_b200/csrc/egs_surfel_math.cuh) for the CPU so that tests without a GPU can compare the exact
// arithmetic the kernels run against the oracle.  Never loaded by the product path.
#include "../../eggfusion_b200/csrc/egs_surfel_math.cuh"
#include "../../eggfusion_b200/csrc/egm_math.cuh"
#include "../../eggfusion_b200/csrc/egt_gn_math.cuh"
#include "../../eggfusion_b200/csrc/egt_qr.cuh"
#include <cmath>
#include <cstring>

static FrameConst make_fc(int W, int H, int D, int M, float tanfovx, float tanfovy, float cx, float cy, float mod,
                          const float* view, const float* proj, const float* campos, const float* bg) {
    FrameConst fc;
    memcpy(fc.view, view, 64);
    memcpy(fc.proj, proj, 64);
    memcpy(fc.campos, campos, 12);
    memcpy(fc.bg, bg, 12);
    fc.tanfovx = tanfovx; fc.tanfovy = tanfovy;
    fc.fy = H / (2.0f * tanfovy);
    fc.fx = W / (2.0f * tanfovx);
    fc.cx = cx; fc.cy = cy; fc.mod = mod;
    fc.W = W; fc.H = H; fc.gx = (W + 15) / 16; fc.gy = (H + 15) / 16; fc.D = D; fc.M = M;
    return fc;
}

extern "C" {

// the product's 6x6 solve (egt_solve_block) on the CPU
int emu_colpiv_qr_solve(const float* A, const float* b, float lm, float* x, int n) {
    return egt_colpiv_qr_solve(A, b, lm, x, n);
}

void emu_surfel_forward(int P, int W, int H, int D, int M, float tanfovx, float tanfovy, float cx, float cy, float mod,
                        const float* view, const float* proj, const float* campos, const float* bg, const float* means,
                        const float* scales, const float* rots, const float* opac, const float* shs,
                        const float* colors, int32_t* radii, uint8_t* active, float* records /*[P][16]*/,
                        float* cov3D /*[P][6]*/, uint8_t* clamped /*[P]*/, int32_t* rect /*[P][4]*/) {
    const FrameConst fc = make_fc(W, H, D, M, tanfovx, tanfovy, cx, cy, mod, view, proj, campos, bg);
    for (int i = 0; i < P; i++) {
        SurfelFwd o;
        memset(&o, 0, sizeof(o));
        const bool use_sh = colors == nullptr;
        surfel_forward(fc, means + 3 * i, scales + 3 * i, rots + 4 * i, opac[i], o);
        if (o.radius > 0) surfel_color(fc, means + 3 * i, use_sh ? shs + (size_t)3 * M * i : colors + 3 * i, use_sh, o);
        radii[i] = o.radius;
        active[i] = (uint8_t)o.active;
        if (o.radius > 0) {
            memcpy(records + 16 * (size_t)i, &o.rec, 64);
            memcpy(cov3D + 6 * (size_t)i, o.cov3D, 24);
            clamped[i] = (uint8_t)o.clamped;
            rect[4 * i] = o.x0; rect[4 * i + 1] = o.y0; rect[4 * i + 2] = o.x1; rect[4 * i + 3] = o.y1;
        }
    }
}

// the sharded projection's conservative footprint bound (surfel_bound_rect): rect[P][4], decided[P]
void emu_surfel_bound_rect(int P, int W, int H, float tanfovx, float tanfovy, float cx, float cy, float mod,
                           const float* view, const float* proj, const float* campos, const float* means,
                           const float* scales, const float* rots, int32_t* rect, uint8_t* decided) {
    const float bg[3] = {0.f, 0.f, 0.f};
    const FrameConst fc = make_fc(W, H, 0, 1, tanfovx, tanfovy, cx, cy, mod, view, proj, campos, bg);
    for (int i = 0; i < P; i++) {
        int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        decided[i] = surfel_bound_rect(fc, means + 3 * i, scales + 3 * i, rots + 4 * i, x0, y0, x1, y1) ? 1 : 0;
        rect[4 * i] = x0; rect[4 * i + 1] = y0; rect[4 * i + 2] = x1; rect[4 * i + 3] = y1;
    }
}

void emu_surfel_backward(int P, int W, int H, int D, int M, float tanfovx, float tanfovy, float cx, float cy, float mod,
                         const float* view, const float* proj, const float* campos, const float* bg, const float* means,
                         const float* scales, const float* rots, const float* shs, const int32_t* radii,
                         const float* cov3D, const uint8_t* clamped, const float* screen_grads /*[P][16]*/,
                         float* d_means, float* d_sh, float* d_scales, float* d_rots, float* d_cov3D) {
    const FrameConst fc = make_fc(W, H, D, M, tanfovx, tanfovy, cx, cy, mod, view, proj, campos, bg);
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        SurfelBwd o;
        const float* g16 = screen_grads + 16 * (size_t)i;
        surfel_backward_geom(fc, means + 3 * i, scales + 3 * i, rots + 4 * i, cov3D + 6 * (size_t)i, g16, o);
        if (shs) {
            const float* mean = means + 3 * i;
            const float dir[3] = {mean[0] - fc.campos[0], mean[1] - fc.campos[1], mean[2] - fc.campos[2]};
            const float gcol[3] = {g16[6], g16[7], g16[8]};
            float add[3];
            float* my_sh = d_sh + (size_t)3 * M * i;
            sh_backward(D, shs + (size_t)3 * M * i, dir, (uint32_t)clamped[i], gcol,
                        [my_sh](int k, int ch, float v) { my_sh[3 * k + ch] = v; }, add);
            o.d_mean[0] += add[0]; o.d_mean[1] += add[1]; o.d_mean[2] += add[2];
        }
        memcpy(d_means + 3 * (size_t)i, o.d_mean, 12);
        memcpy(d_scales + 3 * (size_t)i, o.d_scale, 12);
        memcpy(d_rots + 4 * (size_t)i, o.d_rot, 16);
        memcpy(d_cov3D + 6 * (size_t)i, o.d_cov3D, 24);
    }
}

// ---- mapping glue (egm_math.cuh): the same per-pixel / per-surfel sequence egm_mapping.cu runs -------------------
void emu_cosdist(int n, const float* x1, const float* x2, float up, float* val, float* grad) {
    for (int i = 0; i < n; i++) {
        float d[3] = {0.f, 0.f, 0.f};
        val[i] = egm_cosdist(x1 + 3 * i, x2 + 3 * i, up, d);
        memcpy(grad + 3 * i, d, 12);
    }
}

void emu_adam_geom(int P, int step, const float* lrs /* xyz, opacity, scaling, rotation */, float reg_w, float reg_wn,
                   double nrm2, float* xyz, float* opacity_raw, float* scaling_raw, float* rotation_raw,
                   const float* d_xyz, const float* d_opacity, const float* d_scales, const float* d_rot, float* m_xyz,
                   float* v_xyz, float* m_op, float* v_op, float* m_sc, float* v_sc, float* m_rot, float* v_rot,
                   const float* pos0, const float* normal0, float* g_xyz, float* g_op, float* g_sc, float* g_rot,
                   float* act_o, float* act_s, float* act_r) {
    const double beta1 = 0.9, beta2 = 0.999;
    const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
    EgmAdamConst c;
    c.beta1 = (float)beta1; c.beta2 = (float)beta2;
    c.one_m_beta1 = (float)(1.0 - beta1); c.one_m_beta2 = (float)(1.0 - beta2);
    c.eps = 1e-8f; c.bc2_sqrt = (float)std::sqrt(bc2);
    float nss[4];
    for (int k = 0; k < 4; k++) nss[k] = (float)(-((double)lrs[k] / bc1));
    const float pos_scale = nrm2 > 0.0 ? reg_w / (float)std::sqrt(nrm2) : 0.f;
    for (int i = 0; i < P; i++) {
        EgmSurfel p, g;
        for (int k = 0; k < 3; k++) {
            p.x[k] = xyz[3 * i + k]; p.s[k] = scaling_raw[3 * i + k];
            g.x[k] = d_xyz[3 * i + k]; g.s[k] = d_scales[3 * i + k];
        }
        for (int k = 0; k < 4; k++) { p.q[k] = rotation_raw[4 * i + k]; g.q[k] = d_rot[4 * i + k]; }
        p.o = opacity_raw[i]; g.o = d_opacity[i];
        egm_surfel_raw_grads(p, g, reg_w > 0.f, pos_scale, reg_w * reg_wn / (float)P, pos0 + 3 * i, normal0 + 3 * i);
        memcpy(g_xyz + 3 * i, g.x, 12); memcpy(g_sc + 3 * i, g.s, 12); memcpy(g_rot + 4 * i, g.q, 16); g_op[i] = g.o;
        for (int k = 0; k < 3; k++) {
            xyz[3 * i + k] = egm_adam_update(p.x[k], g.x[k], m_xyz[3 * i + k], v_xyz[3 * i + k], c, nss[0]);
            scaling_raw[3 * i + k] = egm_adam_update(p.s[k], g.s[k], m_sc[3 * i + k], v_sc[3 * i + k], c, nss[2]);
        }
        opacity_raw[i] = egm_adam_update(p.o, g.o, m_op[i], v_op[i], c, nss[1]);
        for (int k = 0; k < 4; k++)
            rotation_raw[4 * i + k] = egm_adam_update(p.q[k], g.q[k], m_rot[4 * i + k], v_rot[4 * i + k], c, nss[3]);
        act_o[i] = egm_sigmoid(opacity_raw[i]);
        for (int k = 0; k < 3; k++) act_s[3 * i + k] = expf(scaling_raw[3 * i + k]);
        EgmRot rot;
        egm_normalize_quat(rotation_raw + 4 * i, rot);
        for (int k = 0; k < 4; k++) act_r[4 * i + k] = egm_nan_to_num(rot.qh[k]);
    }
}

// ---- dense-tracker Gauss-Newton rows (egt_gn_math.cuh): the per-pixel sequence of k_gn_accumulate, sums in double ---
void emu_gn_accumulate(const egt_level* lv, const float* T, float angle_thres_deg, float dist_thres, int use_rgb,
                       double* sums /*[56]*/) {
    for (int i = 0; i < 56; i++) sums[i] = 0.0;
    const float sine_thres = (float)((double)angle_thres_deg * 3.14159265358979323846 / 180.0);
    const long long n = (long long)lv->width * lv->height;
    for (long long p = 0; p < n; p++) {
        GnWarp w;
        egt_gn_warp(*lv, T, p, true, w);
        if (!w.base) continue;
        float J[6], r;
        for (int term = 0; term < (use_rgb ? 2 : 1); term++) {
            const bool valid = term == 0 ? egt_gn_icp_row(*lv, T, p, w, sine_thres, dist_thres, J, r)
                                         : egt_gn_rgb_row(*lv, p, w, J, r);
            if (!valid) continue;
            double* s = sums + 28 * term;
            int k = 0;
            for (int a = 0; a < 6; a++)
                for (int b = a; b < 6; b++) s[k++] += (double)(J[a] * J[b]);
            for (int a = 0; a < 6; a++) s[21 + a] += (double)(J[a] * r);
            s[27] += 1.0;
        }
    }
}
}
