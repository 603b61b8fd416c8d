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

// one pixel of compute_loss: accumulates the partial sums and returns the seeds
struct PixelIn { float ec[3], en[3], ed, rc[3], rn[3], rd; bool m; };
struct PixelOut { float gc[3], gn[3], gd; };
__device__ __forceinline__ void loss_pixel(const PixelIn& p, bool have_depth, bool have_normal, float up_c, float up_d,
                                           float up_n, double (&acc)[4], PixelOut& o) {
    int nans = (p.ed != p.ed) + (p.rd != p.rd);
#pragma unroll
    for (int c = 0; c < 3; c++)
        nans += (p.ec[c] != p.ec[c]) + (p.en[c] != p.en[c]) + (p.rc[c] != p.rc[c]) + (p.rn[c] != p.rn[c]);
    acc[3] += (double)nans;
#pragma unroll
    for (int c = 0; c < 3; c++) { o.gc[c] = 0.f; o.gn[c] = 0.f; }
    o.gd = 0.f;
    if (p.m) {
        // color_loss = |ref - est|[mask].mean()                                        mapper.py:411
        float sc = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float x = p.rc[c] - p.ec[c];
            sc += fabsf(x);
            o.gc[c] = -egm_sign(x) * up_c;
        }
        acc[0] += (double)sc;
        if (have_depth) {   // depth_loss = |ref - est|[mask].mean()                    mapper.py:414-418
            const float x = p.rd - p.ed;
            acc[1] += (double)fabsf(x);
            o.gd = -egm_sign(x) * up_d;
        }
        if (have_normal)    // normal_loss = |1 - cos.clamp|[mask].mean()               mapper.py:421-425
            acc[2] += (double)egm_cosdist(p.rn, p.en, up_n, o.gn);
    }
}

// VEC = 4: a thread owns 4 consecutive pixels -> every access is a 16-byte vector (channel-major float4 per
// channel, pixel-major 3 x float4 per map, masks as one 32-bit word); VEC = 1: generic tail / unaligned fallback.
template <int VEC>
__global__ void __launch_bounds__(256)
k_loss_seed(long long n, const float* __restrict__ est_color, const float* __restrict__ est_depth,
            const float* __restrict__ est_normal, const float* __restrict__ ref_color,
            const float* __restrict__ ref_depth, const float* __restrict__ ref_normal,
            const uint8_t* __restrict__ rgb_mask, const uint8_t* __restrict__ geo_mask, float cw, float dw, float nw,
            float* __restrict__ g_color, float* __restrict__ g_depth, float* __restrict__ g_normal, double* terms,
            const int32_t* __restrict__ tile_mask, int width, int tiles_x) {
    const double cnt = terms[EGM_T_COUNT];
    // mean backward: grad / numel of the indexed tensor ([n,3] colour, [n,1] depth, [n] cosine distance)
    const float up_c = cnt > 0.0 ? (float)((double)cw / (3.0 * cnt)) : 0.f;
    const float up_d = cnt > 0.0 ? (float)((double)dw / cnt) : 0.f;
    const float up_n = cnt > 0.0 ? (float)((double)nw / cnt) : 0.f;
    const bool have_depth = ref_depth != nullptr && dw > 0.f, have_normal = ref_normal != nullptr && nw > 0.f;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // colour, depth, normal, NaN count
    const long long i0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * VEC;
    if (i0 < n) {
        PixelIn p[VEC];
        PixelOut o[VEC];
        if (VEC == 4) {
            union F4 { float4 v; float f[4]; };
            union F12 { float4 v[3]; float f[12]; };
            F4 c[3], nn[3], d, rdv;
            F12 rc, rn;
            rdv.v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                c[k].v = *reinterpret_cast<const float4*>(est_color + k * n + i0);
                nn[k].v = *reinterpret_cast<const float4*>(est_normal + k * n + i0);
                rc.v[k] = reinterpret_cast<const float4*>(ref_color + 3 * i0)[k];
                rn.v[k] = ref_normal ? reinterpret_cast<const float4*>(ref_normal + 3 * i0)[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            d.v = *reinterpret_cast<const float4*>(est_depth + i0);
            if (ref_depth) rdv.v = *reinterpret_cast<const float4*>(ref_depth + i0);
            const uint32_t m0 = *reinterpret_cast<const uint32_t*>(rgb_mask + i0);
            const uint32_t m1 = geo_mask ? *reinterpret_cast<const uint32_t*>(geo_mask + i0) : 0x01010101u;
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    p[j].ec[k] = c[k].f[j]; p[j].en[k] = nn[k].f[j];
                    p[j].rc[k] = rc.f[3 * j + k]; p[j].rn[k] = rn.f[3 * j + k];
                }
                p[j].ed = d.f[j]; p[j].rd = rdv.f[j];
                p[j].m = ((m0 >> (8 * j)) & 0xffu) != 0u && ((m1 >> (8 * j)) & 0xffu) != 0u;
            }
            if (tile_mask) {   // a rank of a tile-sharded frame: only the pixels of its own tiles take part
                const int y = (int)(i0 / width), x = (int)(i0 - (long long)y * width);   // n % 4 == 0 and aligned rows: the 4 pixels share a tile when width % 4 == 0
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int xx = x + j, yy = y + (xx >= width ? 1 : 0), xw = xx >= width ? xx - width : xx;
                    p[j].m = p[j].m && __ldg(tile_mask + (yy >> 4) * tiles_x + (xw >> 4)) != 0;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) loss_pixel(p[j], have_depth, have_normal, up_c, up_d, up_n, acc, o[j]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                *reinterpret_cast<float4*>(g_color + k * n + i0) = make_float4(o[0].gc[k], o[1].gc[k], o[2].gc[k], o[3].gc[k]);
                *reinterpret_cast<float4*>(g_normal + k * n + i0) = make_float4(o[0].gn[k], o[1].gn[k], o[2].gn[k], o[3].gn[k]);
            }
            *reinterpret_cast<float4*>(g_depth + i0) = make_float4(o[0].gd, o[1].gd, o[2].gd, o[3].gd);
        } else {
            const long long i = i0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                p[0].ec[k] = est_color[k * n + i];
                p[0].en[k] = est_normal[k * n + i];
                p[0].rc[k] = ref_color[3 * i + k];
                p[0].rn[k] = ref_normal ? ref_normal[3 * i + k] : 0.f;
            }
            p[0].ed = est_depth[i];
            p[0].rd = ref_depth ? ref_depth[i] : 0.f;
            p[0].m = rgb_mask[i] != 0 && (geo_mask == nullptr || geo_mask[i] != 0);
            if (tile_mask) {
                const int y = (int)(i / width), x = (int)(i - (long long)y * width);
                p[0].m = p[0].m && __ldg(tile_mask + (y >> 4) * tiles_x + (x >> 4)) != 0;
            }
            loss_pixel(p[0], have_depth, have_normal, up_c, up_d, up_n, acc, o[0]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                g_color[k * n + i] = o[0].gc[k];
                g_normal[k * n + i] = o[0].gn[k];
            }
            g_depth[i] = o[0].gd;
        }
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

// [P,3] arrays are staged through shared memory: a CTA of 128 surfels moves 384 consecutive floats per array with
// fully coalesced accesses and each thread then picks its 3 values at stride 3 (conflict-free, gcd(3, 32) = 1);
// reading them straight from global memory at a 12-byte stride per thread ran at 2.0 TB/s (ncu), see profiles/.
#define GEOM_CTA 128
enum { S_XYZ, S_SC, S_DXYZ, S_DSC, S_MX, S_VX, S_MS, S_VS, S_POS0, S_N0, S_SLOTS };

__global__ void __launch_bounds__(GEOM_CTA)
k_adam_geom(int P, GeomArgs a, EgmAdamConst c, float nss_xyz, float nss_opacity, float nss_scaling, float nss_rotation,
            float reg_w, float reg_wn, int step) {
    __shared__ float s3[S_SLOTS][3 * GEOM_CTA];
    const int base = blockIdx.x * GEOM_CTA;
    const int i = base + threadIdx.x;
    const int nv = min(GEOM_CTA, P - base) * 3;   // floats of this CTA's rows
    const bool reg_on = reg_w > 0.f;
    {
        // all 30 loads of a thread are issued before the first shared-memory store (independent, fully unrolled):
        // with a rolled loop every load waited for the previous one's store and the kernel ran latency-bound
        const float* src[S_SLOTS] = {a.xyz, a.scaling_raw, a.d_xyz, a.d_scales, a.m_xyz, a.v_xyz, a.m_scaling,
                                     a.v_scaling, a.pos0, a.normal0};
        float v[S_SLOTS][3];
#pragma unroll
        for (int t = 0; t < S_SLOTS; t++) {
            const bool on = t < S_POS0 || reg_on;
            const float* g = src[t] + 3 * (size_t)base;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int j = threadIdx.x + r * GEOM_CTA;
                v[t][r] = (on && j < nv) ? __ldg(g + j) : 0.f;
            }
        }
#pragma unroll
        for (int t = 0; t < S_SLOTS; t++)
#pragma unroll
            for (int r = 0; r < 3; r++) s3[t][threadIdx.x + r * GEOM_CTA] = v[t][r];
    }
    __syncthreads();
    double acc[2] = {0.0, 0.0};   // sum |1 - cos| at the current parameters, sum (pos0 - xyz_new)^2
    if (i < P) {
        const int l = 3 * threadIdx.x;
        EgmSurfel p, g;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            p.x[k] = s3[S_XYZ][l + k]; p.s[k] = s3[S_SC][l + k];
            g.x[k] = s3[S_DXYZ][l + k]; g.s[k] = s3[S_DSC][l + k];
        }
        const float4 q4 = reinterpret_cast<const float4*>(a.rotation_raw)[i];
        const float4 g4 = reinterpret_cast<const float4*>(a.d_rotations)[i];
        p.q[0] = q4.x; p.q[1] = q4.y; p.q[2] = q4.z; p.q[3] = q4.w;
        g.q[0] = g4.x; g.q[1] = g4.y; g.q[2] = g4.z; g.q[3] = g4.w;
        p.o = a.opacity_raw[i];
        g.o = a.d_opacity[i];

        // ---- activation backward + regulariser (egm_math.cuh)
        float pos0[3] = {0.f, 0.f, 0.f}, n0[3] = {0.f, 0.f, 0.f}, pos_scale = 0.f;
        if (reg_on) {
#pragma unroll
            for (int k = 0; k < 3; k++) { pos0[k] = s3[S_POS0][l + k]; n0[k] = s3[S_N0][l + k]; }
            const double nrm2 = a.reg[step & 1];
            pos_scale = nrm2 > 0.0 ? reg_w / (float)sqrt(nrm2) : 0.f;
        }
        acc[0] = (double)egm_surfel_raw_grads(p, g, reg_on, pos_scale, reg_w * reg_wn / (float)P, pos0, n0);

        // ---- Adam; results go back to the staging slots (S_DSC is reused for the activated scales)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float mm = s3[S_MX][l + k], vv = s3[S_VX][l + k];
            p.x[k] = egm_adam_update(p.x[k], g.x[k], mm, vv, c, nss_xyz);
            s3[S_MX][l + k] = mm; s3[S_VX][l + k] = vv; s3[S_XYZ][l + k] = p.x[k];
            mm = s3[S_MS][l + k]; vv = s3[S_VS][l + k];
            p.s[k] = egm_adam_update(p.s[k], g.s[k], mm, vv, c, nss_scaling);
            s3[S_MS][l + k] = mm; s3[S_VS][l + k] = vv; s3[S_SC][l + k] = p.s[k];
            s3[S_DSC][l + k] = expf(p.s[k]);
            if (reg_on) {
                const float d = pos0[k] - p.x[k];
                acc[1] += (double)(d * d);
            }
        }
        float mm = a.m_opacity[i], vv = a.v_opacity[i];
        const float on = egm_adam_update(p.o, g.o, mm, vv, c, nss_opacity);
        a.m_opacity[i] = mm; a.v_opacity[i] = vv; a.opacity_raw[i] = on;
        a.opacity[i] = egm_sigmoid(on);
        float4 m4 = reinterpret_cast<const float4*>(a.m_rotation)[i], v4 = reinterpret_cast<const float4*>(a.v_rotation)[i];
        float q[4];
        q[0] = egm_adam_update(p.q[0], g.q[0], m4.x, v4.x, c, nss_rotation);
        q[1] = egm_adam_update(p.q[1], g.q[1], m4.y, v4.y, c, nss_rotation);
        q[2] = egm_adam_update(p.q[2], g.q[2], m4.z, v4.z, c, nss_rotation);
        q[3] = egm_adam_update(p.q[3], g.q[3], m4.w, v4.w, c, nss_rotation);
        reinterpret_cast<float4*>(a.m_rotation)[i] = m4;
        reinterpret_cast<float4*>(a.v_rotation)[i] = v4;
        reinterpret_cast<float4*>(a.rotation_raw)[i] = make_float4(q[0], q[1], q[2], q[3]);
        EgmRot rot;
        egm_normalize_quat(q, rot);
        reinterpret_cast<float4*>(a.rotations)[i] = make_float4(egm_nan_to_num(rot.qh[0]), egm_nan_to_num(rot.qh[1]),
                                                                egm_nan_to_num(rot.qh[2]), egm_nan_to_num(rot.qh[3]));
    }
    __syncthreads();
    {
        float* dst[7] = {a.xyz, a.scaling_raw, a.m_xyz, a.v_xyz, a.m_scaling, a.v_scaling, a.scales};
        const int slot[7] = {S_XYZ, S_SC, S_MX, S_VX, S_MS, S_VS, S_DSC};
#pragma unroll
        for (int t = 0; t < 7; t++) {
            float* g = dst[t] + 3 * (size_t)base;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int j = threadIdx.x + r * GEOM_CTA;
                if (j < nv) g[j] = s3[slot[t]][j];
            }
        }
    }
    if (reg_on) {
        // slot 2: sum |1 - cos|; slot ((step + 1) & 1): next norm^2
        __shared__ double s_part[2][GEOM_CTA / 32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const double sred = warp_sum_d(acc[t]);
            if (lane == 0) s_part[t][warp] = sred;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int wv = 0; wv < GEOM_CTA / 32; wv++) { s0 += s_part[0][wv]; s1 += s_part[1][wv]; }
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
    return egm_loss_seed_tiles(height, width, est_color, est_depth, est_normal, ref_color, ref_depth, ref_normal, rgb_mask,
                               geo_mask, nullptr, color_weight, depth_weight, normal_weight, dL_dcolor, dL_ddepth,
                               dL_dnormal, terms, stream);
}

EGS_API int egm_loss_seed_tiles(int32_t height, int32_t width, const float* est_color, const float* est_depth,
                                const float* est_normal, const float* ref_color, const float* ref_depth,
                                const float* ref_normal, const uint8_t* rgb_mask, const uint8_t* geo_mask,
                                const int32_t* tile_mask, float color_weight, float depth_weight, float normal_weight,
                                float* dL_dcolor, float* dL_ddepth, float* dL_dnormal, double* terms, void* stream) {
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
    const uintptr_t align = (uintptr_t)est_color | (uintptr_t)est_depth | (uintptr_t)est_normal | (uintptr_t)ref_color |
                            (uintptr_t)ref_depth | (uintptr_t)ref_normal | (uintptr_t)dL_dcolor | (uintptr_t)dL_ddepth |
                            (uintptr_t)dL_dnormal;
    const uintptr_t malign = (uintptr_t)rgb_mask | (uintptr_t)geo_mask;
    if (n % 4 == 0 && (align & 15) == 0 && (malign & 3) == 0)
        k_loss_seed<4><<<grid_for(n / 4, 256, 1 << 30), 256, 0, s>>>(n, est_color, est_depth, est_normal, ref_color,
                                                                   ref_depth, ref_normal, rgb_mask, geo_mask,
                                                                   color_weight, depth_weight, normal_weight, dL_dcolor,
                                                                   dL_ddepth, dL_dnormal, terms, tile_mask, width,
                                                                   (width + 15) / 16);
    else
        k_loss_seed<1><<<grid_for(n, 256, 1 << 30), 256, 0, s>>>(n, est_color, est_depth, est_normal, ref_color,
                                                               ref_depth, ref_normal, rgb_mask, geo_mask, color_weight,
                                                               depth_weight, normal_weight, dL_dcolor, dL_ddepth,
                                                               dL_dnormal, terms, tile_mask, width, (width + 15) / 16);
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
    k_adam_geom<<<(P + GEOM_CTA - 1) / GEOM_CTA, GEOM_CTA, 0, s>>>(P, a, c, nss(h->lr_xyz), nss(h->lr_opacity), nss(h->lr_scaling),
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
