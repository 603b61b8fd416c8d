// egm_math.cuh -- per-pixel and per-surfel arithmetic of the mapping-iteration glue (include/eggmap.h), written
// __host__ __device__ so that tests/hostemu can run exactly this code on the CPU against the oracle.
//
// Each function restates what PyTorch's autograd computes for the reference's expressions
// (/root/reference/src/core/mapper.py:381-438, gaussian_surfels.py:345-425, core/utils.py:69-92), operation by
// operation; comments name the torch op whose forward / backward formula a line follows.
#pragma once
#include "egs_common.cuh"

// Division and square root.  Device: MUFU-based approximations (div.approx / sqrt.approx, <= 2 ulp): the IEEE
// sequences with their slow-path calls made k_adam_geom issue-bound (5100 SASS instructions, 96 CALLs per surfel;
// 1.75 TB/s of DRAM traffic).  Host (tests/hostemu): exact, like the oracle.  Both are far inside the 1e-4 tolerance.
#if defined(__CUDA_ARCH__)
EGS_HD float egm_div(float a, float b) { return __fdividef(a, b); }
EGS_HD float egm_sqrt(float a) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
#else
EGS_HD float egm_div(float a, float b) { return a / b; }
EGS_HD float egm_sqrt(float a) { return sqrtf(a); }
#endif

// ---- F.cosine_similarity(x1, x2, dim=-1) on 3-vectors + .clamp(-1+1e-6, 1-1e-6) + (1 - .) + abs ----------------
// ATen: x_norm = linalg_vector_norm(x).clone(); x_norm.clamp_min_(eps) under no_grad; cos = sum((x1/x1_norm)*(x2/x2_norm)).
// The clamp happens outside autograd, so the norm's backward still divides by the TRUE norm (0 -> masked to 0).
// Returns |1 - clamp(cos)|; if `up` != 0 adds up * d|1 - clamp(cos)|/d x2 to dx2.
EGS_HD float egm_cosdist(const float x1[3], const float x2[3], float up, float dx2[3]) {
    const float eps = 1e-8f;
    const float n1t = egm_sqrt(x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2]);
    const float n2t = egm_sqrt(x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2]);
    const float n1 = fmaxf(n1t, eps), n2 = fmaxf(n2t, eps);
    const float a0 = egm_div(x1[0], n1), a1 = egm_div(x1[1], n1), a2 = egm_div(x1[2], n1);
    const float b0 = egm_div(x2[0], n2), b1 = egm_div(x2[1], n2), b2 = egm_div(x2[2], n2);
    const float c = a0 * b0 + a1 * b1 + a2 * b2;
    const float lo = (float)(-1 + 1e-6), hi = (float)(1 - 1e-6);
    const float cc = fminf(fmaxf(c, lo), hi);
    const float cd = 1.0f - cc;
    if (up != 0.0f) {
        // abs: grad * sign(cd); rsub: -1; clamp: grad * (lo <= c <= hi)
        const float sg = cd > 0.f ? 1.f : (cd < 0.f ? -1.f : 0.f);
        const float g = (c >= lo && c <= hi) ? -up * sg : 0.0f;
        // mul + sum: d/d(x2/n2) = g * (x1/n1); div: d/dx2 = . / n2, d/dn2 = -sum(. * (x2/n2) / n2)
        const float y0 = g * a0, y1 = g * a1, y2 = g * a2;
        float e0 = egm_div(y0, n2), e1 = egm_div(y1, n2), e2 = egm_div(y2, n2);
        const float dn2 = -(y0 * egm_div(b0, n2) + y1 * egm_div(b1, n2) + y2 * egm_div(b2, n2));
        if (n2t > 0.f) {   // linalg_vector_norm backward: x * (grad / norm), masked where norm == 0
            const float s = egm_div(dn2, n2t);
            e0 += x2[0] * s; e1 += x2[1] * s; e2 += x2[2] * s;
        }
        dx2[0] += e0; dx2[1] += e1; dx2[2] += e2;
    }
    return fabsf(cd);
}

EGS_HD float egm_sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// ---- activations (gaussian_surfels.py:345-425; mapper.py:565-585) ----------------------------------------------
EGS_HD float egm_sigmoid(float x) { return egm_div(1.0f, 1.0f + expf(-x)); }

struct EgmRot {
    float qh[4];   // F.normalize(raw): raw / max(||raw||, 1e-12)
    float nq, den; // ||raw||, clamped denominator
};
EGS_HD void egm_normalize_quat(const float raw[4], EgmRot& r) {
    r.nq = egm_sqrt(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2] + raw[3] * raw[3]);
    r.den = fmaxf(r.nq, 1e-12f);
#pragma unroll
    for (int i = 0; i < 4; i++) r.qh[i] = egm_div(raw[i], r.den);
}
// Mapper.total_params: rotations = torch.nan_to_num(get_rotation, nan=1.0) (+-inf -> +-FLT_MAX)
EGS_HD float egm_nan_to_num(float x) {
    if (x != x) return 1.0f;
    if (x > 3.4028234663852886e38f) return 3.4028234663852886e38f;
    if (x < -3.4028234663852886e38f) return -3.4028234663852886e38f;
    return x;
}
EGS_HD bool egm_finite(float x) { return x == x && x <= 3.4028234663852886e38f && x >= -3.4028234663852886e38f; }

// F.normalize backward: given dL/dqh (already summed over all consumers of qh) -> dL/draw
EGS_HD void egm_normalize_quat_bwd(const float raw[4], const EgmRot& r, const float dqh[4], float draw[4]) {
    float dden = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        draw[i] = egm_div(dqh[i], r.den);
        dden -= dqh[i] * egm_div(r.qh[i], r.den);
    }
    // clamp_min backward: grad * (norm >= eps); norm backward: x * (grad / norm), masked where norm == 0
    if (r.nq >= 1e-12f && r.nq > 0.f) {
        const float s = egm_div(dden, r.nq);
#pragma unroll
        for (int i = 0; i < 4; i++) draw[i] += raw[i] * s;
    }
}

// GaussianSurfels.get_normal: column argmin(scales) of build_rotation(get_rotation), divided by (its norm + 1e-8).
struct EgmNormal {
    float q[4];    // build_rotation's own re-normalised quaternion (r, x, y, z)
    float nb;      // its norm sqrt(sum qh^2)
    float v[3];    // selected column of R
    float mag;     // ||v||
    int k;
};
EGS_HD int egm_argmin3(float a, float b, float c) {   // torch.argmin: first minimal index
    int k = 0;
    float m = a;
    if (b < m) { m = b; k = 1; }
    if (c < m) { k = 2; }
    return k;
}
EGS_HD void egm_get_normal(const float qh[4], int k, EgmNormal& s, float n[3]) {
    s.k = k;
    s.nb = egm_sqrt(qh[0] * qh[0] + qh[1] * qh[1] + qh[2] * qh[2] + qh[3] * qh[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) s.q[i] = egm_div(qh[i], s.nb);
    const float r = s.q[0], x = s.q[1], y = s.q[2], z = s.q[3];
    if (k == 0) {
        s.v[0] = 1 - 2 * (y * y + z * z); s.v[1] = 2 * (x * y + r * z); s.v[2] = 2 * (x * z - r * y);
    } else if (k == 1) {
        s.v[0] = 2 * (x * y - r * z); s.v[1] = 1 - 2 * (x * x + z * z); s.v[2] = 2 * (y * z + r * x);
    } else {
        s.v[0] = 2 * (x * z + r * y); s.v[1] = 2 * (y * z - r * x); s.v[2] = 1 - 2 * (x * x + y * y);
    }
    s.mag = egm_sqrt(s.v[0] * s.v[0] + s.v[1] * s.v[1] + s.v[2] * s.v[2]);
    const float d = s.mag + 1e-8f;
    n[0] = egm_div(s.v[0], d); n[1] = egm_div(s.v[1], d); n[2] = egm_div(s.v[2], d);
}
// dL/dn -> adds dL/dqh (gradient w.r.t. the F.normalize output that build_rotation received)
EGS_HD void egm_get_normal_bwd(const float qh[4], const EgmNormal& s, const float dn[3], float dqh[4]) {
    const float d = s.mag + 1e-8f;
    float dv[3];
    float dmag = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        dv[i] = egm_div(dn[i], d);
        dmag -= dn[i] * egm_div(egm_div(s.v[i], d), d);
    }
    if (s.mag > 0.f) {
        const float t = egm_div(dmag, s.mag);
#pragma unroll
        for (int i = 0; i < 3; i++) dv[i] += s.v[i] * t;
    }
    const float r = s.q[0], x = s.q[1], y = s.q[2], z = s.q[3];
    float dq[4];   // (dr, dx, dy, dz)
    if (s.k == 0) {
        dq[0] = 2 * z * dv[1] - 2 * y * dv[2];
        dq[1] = 2 * y * dv[1] + 2 * z * dv[2];
        dq[2] = -4 * y * dv[0] + 2 * x * dv[1] - 2 * r * dv[2];
        dq[3] = -4 * z * dv[0] + 2 * r * dv[1] + 2 * x * dv[2];
    } else if (s.k == 1) {
        dq[0] = -2 * z * dv[0] + 2 * x * dv[2];
        dq[1] = 2 * y * dv[0] - 4 * x * dv[1] + 2 * r * dv[2];
        dq[2] = 2 * x * dv[0] + 2 * z * dv[2];
        dq[3] = -2 * r * dv[0] - 4 * z * dv[1] + 2 * y * dv[2];
    } else {
        dq[0] = 2 * y * dv[0] - 2 * x * dv[1];
        dq[1] = 2 * z * dv[0] - 2 * r * dv[1] - 4 * x * dv[2];
        dq[2] = 2 * r * dv[0] + 2 * z * dv[1] - 4 * y * dv[2];
        dq[3] = 2 * x * dv[0] + 2 * y * dv[1];
    }
    // q = qh / nb with nb = sqrt(sum qh^2)
    float dnb = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) dnb -= dq[i] * egm_div(s.q[i], s.nb);
#pragma unroll
    for (int i = 0; i < 4; i++) dqh[i] += egm_div(dq[i], s.nb) + dnb * egm_div(qh[i], s.nb);
}

// ---- one surfel: gradients w.r.t. the ACTIVATED parameters -> gradients w.r.t. the raw parameters (+ regulariser) ----
struct EgmSurfel {
    float x[3], o, s[3], q[4];   // xyz, opacity, scaling, rotation: raw values, or gradients in the same slots
};
// g in: dL/d(xyz, sigmoid(o), exp(s), nan_to_num(normalize(q))) as egs_backward_surfels writes them;
// g out: dL/d(xyz, o, s, q) including the regulariser of mapper.py:427-435 when reg_n_up / reg_pos_scale are non-zero.
//   reg_pos_scale = reg_weight / ||pos0 - xyz||_F (0 when the norm is 0: torch.norm backward masks it)
//   reg_n_up      = reg_weight * reg_weight_n / P
// Returns |1 - clamp(cos(normal0, get_normal))| of this surfel (0 when the regulariser is off).
EGS_HD float egm_surfel_raw_grads(const EgmSurfel& p, EgmSurfel& g, bool reg_on, float reg_pos_scale, float reg_n_up,
                                  const float pos0[3], const float normal0[3]) {
    // sigmoid backward: g * (1 - y) * y;  exp backward: g * y;  nan_to_num backward: g * isfinite(input)
    const float so = egm_sigmoid(p.o);
    g.o = g.o * (1.0f - so) * so;
    float es[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { es[k] = expf(p.s[k]); g.s[k] = g.s[k] * es[k]; }
    EgmRot rot;
    egm_normalize_quat(p.q, rot);
    float dqh[4];
#pragma unroll
    for (int k = 0; k < 4; k++) dqh[k] = egm_finite(rot.qh[k]) ? g.q[k] : 0.f;
    float cd = 0.f;
    if (reg_on) {
#pragma unroll
        for (int k = 0; k < 3; k++) g.x[k] += (p.x[k] - pos0[k]) * reg_pos_scale;
        EgmNormal ns;
        float n[3], dn[3] = {0.f, 0.f, 0.f};
        egm_get_normal(rot.qh, egm_argmin3(es[0], es[1], es[2]), ns, n);
        cd = egm_cosdist(normal0, n, reg_n_up, dn);
        egm_get_normal_bwd(rot.qh, ns, dn, dqh);
    }
    egm_normalize_quat_bwd(p.q, rot, dqh, g.q);
    return cd;
}

// ---- torch.optim.Adam (single update of one element) ------------------------------------------------------------
struct EgmAdamConst {
    float beta1, beta2, one_m_beta1, one_m_beta2, eps, bc2_sqrt;
};
// neg_step_size = -(lr / (1 - beta1^t)), bc2_sqrt = sqrt(1 - beta2^t) (computed in double on the host, like torch)
EGS_HD float egm_adam_update(float p, float g, float& m, float& v, const EgmAdamConst& c, float neg_step_size) {
    // explicit roundings (the contraction nvcc applies to torch's kernels), so that every kernel this is inlined into --
    // k_adam_sh, k_adam_geom, the fused k_surfel_backward -- produces the same bits
    m = f_fma(c.one_m_beta1, f_sub(g, m), m);                       // exp_avg.lerp_(grad, 1 - beta1), weight < 0.5 branch
    v = f_fma(f_mul(c.one_m_beta2, g), g, f_mul(v, c.beta2));       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = f_add(egm_div(egm_sqrt(v), c.bc2_sqrt), c.eps);
    return f_fma(neg_step_size, egm_div(m, denom), p);              // param.addcdiv_(exp_avg, denom, value=-step_size)
}
