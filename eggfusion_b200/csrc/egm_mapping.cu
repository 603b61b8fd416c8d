// egm_mapping.cu -- the mapping-iteration glue around the rasterizer (include/eggmap.h):
// fused loss + gradient seeds, fused activation-backward + regulariser + Adam + re-activation.
//
// Reference: Mapper.frame_batch_optimization / compute_loss (/root/reference/src/core/mapper.py:336-368,381-444),
// GaussianSurfels activations and parameter groups (gaussian_surfels.py:134-150,345-425), torch.optim.Adam.
// All kernels are HBM streaming passes (no contraction): coalesced loads, one pass over each array, block
// reductions in shuffles + one double atomic per block and term.
//
//   k_mask_count   2 B/pixel            -> number of masked pixels (the means' denominators)
//   k_loss_seed    reads 7+7 floats + 2 B, writes 7 floats per pixel (62 B/pixel)
//   k_adam_sh      elementwise over the P*M*3 SH coefficients: reads g, p, m, v, writes p, m, v (28 B/coefficient)
//   k_adam_geom    per surfel: 11 raw values + 11 gradients + 22 state + anchors in, 11 raw + 22 state + 8
//                  activated values out (~ 330 B/surfel)
#include "egm_math.cuh"
#include "../../include/eggmap.h"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// sums `NV` per-thread doubles over the block and adds them to dst[0..NV) with one atomic per value
template <int NV>
__device__ __forceinline__ void block_accumulate(const double (&v)[NV], double* dst) {
    __shared__ double s_part[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double s = warp_sum_d(v[i]);
        if (lane == 0) s_part[i][warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double s = lane < nw ? s_part[i][lane] : 0.0;
            s = warp_sum_d(s);
            if (lane == 0 && s != 0.0) atomicAdd(dst + i, s);
        }
    }
}

__global__ void __launch_bounds__(256)
k_mask_count(long long n, const uint8_t* __restrict__ rgb_mask, const uint8_t* __restrict__ geo_mask, double* terms) {
    double c[1] = {0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool m = rgb_mask[i] != 0 && (geo_mask == nullptr || geo_mask[i] != 0);
        c[0] += m ? 1.0 : 0.0;
    }
    block_accumulate<1>(c, terms + EGM_T_COUNT);
}

__global__ void __launch_bounds__(256)
k_loss_seed(long long n, const float* __restrict__ est_color, const float* __restrict__ est_depth,
            const float* __restrict__ est_normal, const float* __restrict__ ref_color,
            const float* __restrict__ ref_depth, const float* __restrict__ ref_normal,
            const uint8_t* __restrict__ rgb_mask, const uint8_t* __restrict__ geo_mask, float cw, float dw, float nw,
            float* __restrict__ g_color, float* __restrict__ g_depth, float* __restrict__ g_normal, double* terms) {
    const double cnt = terms[EGM_T_COUNT];
    // mean backward: grad / numel of the indexed tensor ([n,3] colour, [n,1] depth, [n] cosine distance)
    const float up_c = cnt > 0.0 ? (float)((double)cw / (3.0 * cnt)) : 0.f;
    const float up_d = cnt > 0.0 ? (float)((double)dw / cnt) : 0.f;
    const float up_n = cnt > 0.0 ? (float)((double)nw / cnt) : 0.f;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // colour, depth, normal, NaN count
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool m = rgb_mask[i] != 0 && (geo_mask == nullptr || geo_mask[i] != 0);
        float ec[3], en[3], rc[3], rn[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            ec[c] = est_color[c * n + i];
            en[c] = est_normal[c * n + i];
            rc[c] = ref_color[3 * i + c];
            if (ref_normal) rn[c] = ref_normal[3 * i + c];
        }
        const float ed = est_depth[i];
        const float rd = ref_depth ? ref_depth[i] : 0.f;
        int nans = (ed != ed) + (rd != rd);
#pragma unroll
        for (int c = 0; c < 3; c++) nans += (ec[c] != ec[c]) + (en[c] != en[c]) + (rc[c] != rc[c]) + (rn[c] != rn[c]);
        acc[3] += (double)nans;
        float gc[3] = {0.f, 0.f, 0.f}, gn[3] = {0.f, 0.f, 0.f}, gd = 0.f;
        if (m) {
            // color_loss = |ref - est|[mask].mean()                                        mapper.py:411
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float x = rc[c] - ec[c];
                acc[0] += (double)fabsf(x);
                gc[c] = -egm_sign(x) * up_c;
            }
            if (ref_depth && dw > 0.f) {   // depth_loss = |ref - est|[mask].mean()         mapper.py:414-418
                const float x = rd - ed;
                acc[1] += (double)fabsf(x);
                gd = -egm_sign(x) * up_d;
            }
            if (ref_normal && nw > 0.f)    // normal_loss = |1 - cos.clamp|[mask].mean()    mapper.py:421-425
                acc[2] += (double)egm_cosdist(rn, en, up_n, gn);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            g_color[c * n + i] = gc[c];
            g_normal[c * n + i] = gn[c];
        }
        g_depth[i] = gd;
    }
    block_accumulate<4>(acc, terms + EGM_T_COLOR);
}

__global__ void k_loss_total(const double* terms, const double* reg, int step, int P, float cw, float dw, float nw,
                             float rw, float rwn, int have_depth, int have_normal, float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = terms[EGM_T_COUNT];
    const float color = n > 0.0 ? (float)(terms[EGM_T_COLOR] / (3.0 * n)) : __int_as_float(0x7fc00000);
    const float depth = (have_depth && dw > 0.f && n > 0.0) ? (float)(terms[EGM_T_DEPTH] / n) : 0.f;
    const float normal = (have_normal && nw > 0.f && n > 0.0) ? (float)(terms[EGM_T_NORMAL] / n) : 0.f;
    float regl = 0.f;
    if (reg && rw > 0.f && P > 0) {
        const float pos = (float)sqrt(reg[step & 1]);
        regl = pos + rwn * (float)(reg[2] / (double)P);   // reg_position.mean() + reg_weight_n * reg_normal.abs().mean()
    }
    out[1] = color; out[2] = depth; out[3] = normal; out[4] = regl;
    out[0] = cw * color + dw * depth + nw * normal + rw * regl;   // mapper.py:438
}

// ---- Adam over the SH block: identity activation, lr by row (row 0 = f_dc, the others f_rest) --------------------
template <typename VT>
__global__ void __launch_bounds__(256)
k_adam_sh(long long count, int row_elems, const float* __restrict__ g, float* __restrict__ p, float* __restrict__ m,
          float* __restrict__ v, EgmAdamConst c, float nss_dc, float nss_rest) {
    constexpr int VN = sizeof(VT) / 4;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= count) return;
    union U { VT vec; float f[VN]; };
    U ug, up, um, uv;
    ug.vec = reinterpret_cast<const VT*>(g)[i];
    up.vec = reinterpret_cast<const VT*>(p)[i];
    um.vec = reinterpret_cast<const VT*>(m)[i];
    uv.vec = reinterpret_cast<const VT*>(v)[i];
    const int col0 = (int)((i * VN) % row_elems);
#pragma unroll
    for (int k = 0; k < VN; k++) up.f[k] = egm_adam_update(up.f[k], ug.f[k], um.f[k], uv.f[k], c, (col0 + k) < 3 ? nss_dc : nss_rest);
    reinterpret_cast<VT*>(p)[i] = up.vec;
    reinterpret_cast<VT*>(m)[i] = um.vec;
    reinterpret_cast<VT*>(v)[i] = uv.vec;
}

struct GeomArgs {
    float *xyz, *opacity_raw, *scaling_raw, *rotation_raw;
    const float *d_xyz, *d_opacity, *d_scales, *d_rotations;
    float *m_xyz, *v_xyz, *m_opacity, *v_opacity, *m_scaling, *v_scaling, *m_rotation, *v_rotation;
    const float *pos0, *normal0;
    double* reg;
    float *opacity, *scales, *rotations;
};

__global__ void __launch_bounds__(128)
k_adam_geom(int P, GeomArgs a, EgmAdamConst c, float nss_xyz, float nss_opacity, float nss_scaling, float nss_rotation,
            float reg_w, float reg_wn, int step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[2] = {0.0, 0.0};   // sum |1 - cos| at the current parameters, sum (pos0 - xyz_new)^2
    if (i < P) {
        EgmSurfel p, g;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            p.x[k] = a.xyz[3 * i + k]; p.s[k] = a.scaling_raw[3 * i + k];
            g.x[k] = a.d_xyz[3 * i + k]; g.s[k] = a.d_scales[3 * i + k];
        }
        const float4 q4 = reinterpret_cast<const float4*>(a.rotation_raw)[i];
        const float4 g4 = reinterpret_cast<const float4*>(a.d_rotations)[i];
        p.q[0] = q4.x; p.q[1] = q4.y; p.q[2] = q4.z; p.q[3] = q4.w;
        g.q[0] = g4.x; g.q[1] = g4.y; g.q[2] = g4.z; g.q[3] = g4.w;
        p.o = a.opacity_raw[i];
        g.o = a.d_opacity[i];

        // ---- activation backward + regulariser (egm_math.cuh)
        float pos0[3] = {0.f, 0.f, 0.f}, n0[3] = {0.f, 0.f, 0.f}, pos_scale = 0.f;
        if (reg_w > 0.f) {
#pragma unroll
            for (int k = 0; k < 3; k++) { pos0[k] = a.pos0[3 * i + k]; n0[k] = a.normal0[3 * i + k]; }
            const double nrm2 = a.reg[step & 1];
            pos_scale = nrm2 > 0.0 ? reg_w / (float)sqrt(nrm2) : 0.f;
        }
        acc[0] = (double)egm_surfel_raw_grads(p, g, reg_w > 0.f, pos_scale, reg_w * reg_wn / (float)P, pos0, n0);
        float* x = p.x; float* s = p.s; float* q = p.q;
        const float* gx = g.x; const float* gs = g.s; const float* dq = g.q;
        const float o = p.o, go = g.o;
        EgmRot rot;

        // ---- Adam
        float mm, vv;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            mm = a.m_xyz[3 * i + k]; vv = a.v_xyz[3 * i + k];
            x[k] = egm_adam_update(x[k], gx[k], mm, vv, c, nss_xyz);
            a.m_xyz[3 * i + k] = mm; a.v_xyz[3 * i + k] = vv; a.xyz[3 * i + k] = x[k];
            mm = a.m_scaling[3 * i + k]; vv = a.v_scaling[3 * i + k];
            s[k] = egm_adam_update(s[k], gs[k], mm, vv, c, nss_scaling);
            a.m_scaling[3 * i + k] = mm; a.v_scaling[3 * i + k] = vv; a.scaling_raw[3 * i + k] = s[k];
        }
        mm = a.m_opacity[i]; vv = a.v_opacity[i];
        const float on = egm_adam_update(o, go, mm, vv, c, nss_opacity);
        a.m_opacity[i] = mm; a.v_opacity[i] = vv; a.opacity_raw[i] = on;
        float4 m4 = reinterpret_cast<const float4*>(a.m_rotation)[i], v4 = reinterpret_cast<const float4*>(a.v_rotation)[i];
        q[0] = egm_adam_update(q[0], dq[0], m4.x, v4.x, c, nss_rotation);
        q[1] = egm_adam_update(q[1], dq[1], m4.y, v4.y, c, nss_rotation);
        q[2] = egm_adam_update(q[2], dq[2], m4.z, v4.z, c, nss_rotation);
        q[3] = egm_adam_update(q[3], dq[3], m4.w, v4.w, c, nss_rotation);
        reinterpret_cast<float4*>(a.m_rotation)[i] = m4;
        reinterpret_cast<float4*>(a.v_rotation)[i] = v4;
        reinterpret_cast<float4*>(a.rotation_raw)[i] = make_float4(q[0], q[1], q[2], q[3]);

        // ---- activations for the next forward
        a.opacity[i] = egm_sigmoid(on);
#pragma unroll
        for (int k = 0; k < 3; k++) a.scales[3 * i + k] = expf(s[k]);
        egm_normalize_quat(q, rot);
        reinterpret_cast<float4*>(a.rotations)[i] = make_float4(egm_nan_to_num(rot.qh[0]), egm_nan_to_num(rot.qh[1]),
                                                                egm_nan_to_num(rot.qh[2]), egm_nan_to_num(rot.qh[3]));
        if (reg_w > 0.f) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float d = pos0[k] - x[k];
                acc[1] += (double)(d * d);
            }
        }
    }
    if (reg_w > 0.f) {
        const double out[2] = {acc[0], acc[1]};
        // slot 2: sum |1 - cos|; slot ((step + 1) & 1): next norm^2
        __shared__ double s_part[2][4];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const double sred = warp_sum_d(out[t]);
            if (lane == 0) s_part[t][warp] = sred;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double s0 = s_part[0][0] + s_part[0][1] + s_part[0][2] + s_part[0][3];
            const double s1 = s_part[1][0] + s_part[1][1] + s_part[1][2] + s_part[1][3];
            if (s0 != 0.0) atomicAdd(a.reg + 2, s0);
            if (s1 != 0.0) atomicAdd(a.reg + ((step + 1) & 1), s1);
        }
    }
}

__global__ void __launch_bounds__(128)
k_activate(int P, const float* __restrict__ opacity_raw, const float* __restrict__ scaling_raw,
           const float* __restrict__ rotation_raw, float* __restrict__ opacity, float* __restrict__ scales,
           float* __restrict__ rotations, float* __restrict__ normals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float es[3];
#pragma unroll
    for (int k = 0; k < 3; k++) es[k] = expf(scaling_raw[3 * i + k]);
    const float4 q4 = reinterpret_cast<const float4*>(rotation_raw)[i];
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    EgmRot rot;
    egm_normalize_quat(q, rot);
    if (opacity) opacity[i] = egm_sigmoid(opacity_raw[i]);
    if (scales) {
#pragma unroll
        for (int k = 0; k < 3; k++) scales[3 * i + k] = es[k];
    }
    if (rotations)
        reinterpret_cast<float4*>(rotations)[i] = make_float4(egm_nan_to_num(rot.qh[0]), egm_nan_to_num(rot.qh[1]),
                                                              egm_nan_to_num(rot.qh[2]), egm_nan_to_num(rot.qh[3]));
    if (normals) {
        EgmNormal ns;
        float n[3];
        egm_get_normal(rot.qh, egm_argmin3(es[0], es[1], es[2]), ns, n);
#pragma unroll
        for (int k = 0; k < 3; k++) normals[3 * i + k] = n[k];
    }
}

#define EGM_TRY(expr)                              \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

inline int grid_for(long long n, int block, int cap) {
    long long g = (n + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
} // namespace

extern "C" {

EGS_API int egm_loss_seed(int32_t height, int32_t width, const float* est_color, const float* est_depth,
                          const float* est_normal, const float* ref_color, const float* ref_depth,
                          const float* ref_normal, const uint8_t* rgb_mask, const uint8_t* geo_mask,
                          float color_weight, float depth_weight, float normal_weight, float* dL_dcolor,
                          float* dL_ddepth, float* dL_dnormal, double* terms, void* stream) {
    if (height <= 0 || width <= 0) return EGS_E_BADARG;
    if (!est_color || !est_depth || !est_normal || !ref_color || !rgb_mask || !dL_dcolor || !dL_ddepth || !dL_dnormal ||
        !terms)
        return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = (long long)height * width;
    EGM_TRY(cudaMemsetAsync(terms, 0, sizeof(double) * EGM_TERMS, s));
    // 148 SMs x 8 resident 256-thread CTAs; grid-stride beyond that
    k_mask_count<<<grid_for(n, 256 * 16, 148 * 8), 256, 0, s>>>(n, rgb_mask, geo_mask, terms);
    EGM_TRY(cudaGetLastError());
    k_loss_seed<<<grid_for(n, 256, 148 * 8 * 4), 256, 0, s>>>(n, est_color, est_depth, est_normal, ref_color, ref_depth,
                                                             ref_normal, rgb_mask, geo_mask, color_weight, depth_weight,
                                                             normal_weight, dL_dcolor, dL_ddepth, dL_dnormal, terms);
    EGM_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egm_adam_step(int32_t P, int32_t sh_coeffs, const egm_adam* h, float* xyz, float* shs, float* opacity_raw,
                          float* scaling_raw, float* rotation_raw, const float* d_xyz, const float* d_shs,
                          const float* d_opacity, const float* d_scales, const float* d_rotations, float* m_xyz,
                          float* v_xyz, float* m_shs, float* v_shs, float* m_opacity, float* v_opacity,
                          float* m_scaling, float* v_scaling, float* m_rotation, float* v_rotation, const float* pos0,
                          const float* normal0, double* reg, float* opacity, float* scales, float* rotations,
                          void* stream) {
    if (P < 0 || sh_coeffs < 0 || !h || h->step < 1) return EGS_E_BADARG;
    if (P == 0) return 0;
    if (!xyz || !opacity_raw || !scaling_raw || !rotation_raw || !d_xyz || !d_opacity || !d_scales || !d_rotations ||
        !m_xyz || !v_xyz || !m_opacity || !v_opacity || !m_scaling || !v_scaling || !m_rotation || !v_rotation ||
        !opacity || !scales || !rotations)
        return EGS_E_BADARG;
    if (sh_coeffs > 0 && (!shs || !d_shs || !m_shs || !v_shs)) return EGS_E_BADARG;
    if (h->reg_weight > 0.f && (!pos0 || !normal0 || !reg)) return EGS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    // torch computes the bias corrections and the step size in double on the host (optim/adam.py)
    const double bc1 = 1.0 - pow(h->beta1, (double)h->step);
    const double bc2 = 1.0 - pow(h->beta2, (double)h->step);
    EgmAdamConst c;
    c.beta1 = (float)h->beta1; c.beta2 = (float)h->beta2;
    c.one_m_beta1 = (float)(1.0 - h->beta1); c.one_m_beta2 = (float)(1.0 - h->beta2);
    c.eps = (float)h->eps; c.bc2_sqrt = (float)sqrt(bc2);
    auto nss = [&](float lr) { return (float)(-((double)lr / bc1)); };
    if (h->reg_weight > 0.f) {
        // zero the two output slots of the regulariser state (slot 2 and the next norm^2)
        EGM_TRY(cudaMemsetAsync(reg + 2, 0, sizeof(double), s));
        EGM_TRY(cudaMemsetAsync(reg + ((h->step + 1) & 1), 0, sizeof(double), s));
    }
    if (sh_coeffs > 0) {
        const long long elems = (long long)P * sh_coeffs * 3;
        const int row = sh_coeffs * 3;
        if (row % 4 == 0 && (((uintptr_t)shs | (uintptr_t)d_shs | (uintptr_t)m_shs | (uintptr_t)v_shs) & 15) == 0) {
            const long long cnt = elems / 4;
            k_adam_sh<float4><<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(cnt, row, d_shs, shs, m_shs, v_shs, c,
                                                                          nss(h->lr_f_dc), nss(h->lr_f_rest));
        } else {
            k_adam_sh<float><<<(unsigned)((elems + 255) / 256), 256, 0, s>>>(elems, row, d_shs, shs, m_shs, v_shs, c,
                                                                           nss(h->lr_f_dc), nss(h->lr_f_rest));
        }
        EGM_TRY(cudaGetLastError());
    }
    GeomArgs a{xyz, opacity_raw, scaling_raw, rotation_raw, d_xyz, d_opacity, d_scales, d_rotations, m_xyz, v_xyz,
               m_opacity, v_opacity, m_scaling, v_scaling, m_rotation, v_rotation, pos0, normal0, reg, opacity, scales,
               rotations};
    k_adam_geom<<<(P + 127) / 128, 128, 0, s>>>(P, a, c, nss(h->lr_xyz), nss(h->lr_opacity), nss(h->lr_scaling),
                                                nss(h->lr_rotation), h->reg_weight, h->reg_weight_n, h->step);
    EGM_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egm_activate(int32_t P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                         float* opacity, float* scales, float* rotations, float* normals, void* stream) {
    if (P < 0) return EGS_E_BADARG;
    if (P == 0) return 0;
    if (!scaling_raw || !rotation_raw || (opacity && !opacity_raw)) return EGS_E_BADARG;
    k_activate<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, opacity_raw, scaling_raw, rotation_raw, opacity,
                                                                  scales, rotations, normals);
    EGM_TRY(cudaGetLastError());
    return 0;
}

EGS_API int egm_loss_total(const double* terms, const double* reg, int32_t step, int32_t P, float color_weight,
                           float depth_weight, float normal_weight, float reg_weight, float reg_weight_n,
                           int32_t have_depth, int32_t have_normal, float* out, void* stream) {
    if (!terms || !out) return EGS_E_BADARG;
    k_loss_total<<<1, 32, 0, (cudaStream_t)stream>>>(terms, reg, step, P, color_weight, depth_weight, normal_weight,
                                                     reg_weight, reg_weight_n, have_depth, have_normal, out);
    EGM_TRY(cudaGetLastError());
    return 0;
}

} // extern "C"
