// egs_fusion.cu -- once-per-frame surfel fusion kernels.
//
// Replaces projectSurfelsToFrame and preprocessSurfel of the reference
// (/root/reference/submodules/diff-gaussian-surfels/fuse_surfels.cu:475-536 and :214-394).
//   * k_project_surfels: the reference does atomicMin on the float depth and then a plain store of the surfel id
//     when it lowered the value -- a race that can leave a stale id.  Here one 64-bit atomicMin on
//     (depth bits << 32 | id) per footprint pixel makes the outcome deterministic: smallest depth, ties to the
//     smallest id (what the reference produces whenever its race does not bite).  k_unpack_zbuffer then writes the
//     int32 index map / float depth buffer the Python API returns.
//   * k_fuse_surfels: in-place information-filter update of position / normal; no device synchronisation,
//     launched on the caller's stream (the reference calls cudaDeviceSynchronize()).
#include "egs_surfel_math.cuh"

struct FuseCam {
    float view[16], proj[16];
    float cx, cy;
    int wd, ht;
};

__device__ __forceinline__ void load_fuse_cam(FuseCam& c, const float* view, const float* proj, const float* intr,
                                              int wd, int ht) {
    const int t = threadIdx.x;
    if (t < 16) c.view[t] = __ldg(view + t);
    else if (t < 32) c.proj[t - 16] = __ldg(proj + t - 16);
    else if (t == 32) { c.cx = __ldg(intr + 2); c.cy = __ldg(intr + 3); c.wd = wd; c.ht = ht; }
}

// projection + frustum + back-face tests shared by both kernels (same arithmetic as the rasterizer's per-surfel stage)
__device__ __forceinline__ bool fuse_common(const FuseCam& c, const float* p, const float* q, float pc[3],
                                            float coord[2], Mat3& R, float zaxis[3]) {
    const float hx = xf_affine(c.proj, 0, p[0], p[1], p[2]), hy = xf_affine(c.proj, 1, p[0], p[1], p[2]);
    const float hw = xf_affine(c.proj, 3, p[0], p[1], p[2]);
    const float pw = f_rcp(f_add(hw, 0.0000001f));
#pragma unroll
    for (int r = 0; r < 3; r++) pc[r] = xf_affine(c.view, r, p[0], p[1], p[2]);
    coord[0] = (float)fma((double)f_mul(f_mul(hx, pw), (float)c.wd), 0.5, (double)c.cx);
    coord[1] = (float)fma((double)f_mul(f_mul(hy, pw), (float)c.ht), 0.5, (double)c.cy);
    const float e = 0.05f, e1 = f_add(1.0f, e);
    const float x0 = f_mul((float)(-c.wd), e), x1 = f_mul((float)c.wd, e1);
    const float y0 = f_mul((float)(-c.ht), e), y1 = f_mul((float)c.ht, e1);
    if (pc[2] < 0.f || coord[0] < x0 || coord[0] >= x1 || coord[1] < y0 || coord[1] >= y1) return false;
    R = quat_to_rot(q[0], q[1], q[2], q[3]);
#pragma unroll
    for (int r = 0; r < 3; r++) zaxis[r] = xf_linear(c.view, r, R.m[0][2], R.m[1][2], R.m[2][2]);
    const float facing = f_dot3(pc[0], zaxis[0], pc[1], zaxis[1], pc[2], zaxis[2]);
    return !((double)facing > -0.00001);
}

__global__ void __launch_bounds__(256)
k_project_surfels(int P, int ht, int wd, const float* __restrict__ points, const float* __restrict__ rotations,
                  const uint8_t* __restrict__ stable, const float* __restrict__ intr, const float* __restrict__ view,
                  const float* __restrict__ proj, unsigned long long* __restrict__ zbuf) {
    __shared__ FuseCam c;
    load_fuse_cam(c, view, proj, intr, wd, ht);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || !stable[i]) return;
    const float p[3] = {points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]};
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(rotations) + i);
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float pc[3], coord[2], z[3];
    Mat3 R;
    if (!fuse_common(c, p, q, pc, coord, R, z)) return;
    const int x = (int)coord[0], y = (int)coord[1];
    if (x < 0 || x >= wd || y < 0 || y >= ht) return;
    const unsigned long long key = ((unsigned long long)__float_as_uint(pc[2]) << 32) | (unsigned)i;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            const int nx = x + dx, ny = y + dy;
            if (nx < 0 || nx >= wd || ny < 0 || ny >= ht) continue;
            atomicMin(zbuf + (size_t)ny * wd + nx, key);
        }
}

__global__ void k_init_zbuffer(size_t n, unsigned long long* __restrict__ zbuf) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) zbuf[k] = 0x7f800000ffffffffull; // +inf depth, no surfel
}

__global__ void k_unpack_zbuffer(size_t n, const unsigned long long* __restrict__ zbuf, int32_t* __restrict__ index_map,
                                 float* __restrict__ depth) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned long long v = zbuf[k];
    const uint32_t id = (uint32_t)v;
    if (id == 0xffffffffu) return; // untouched pixels keep the caller's -1 / +inf
    // keep an existing nearer entry of the caller's buffers (the reference accumulates into them)
    const float d = __uint_as_float((uint32_t)(v >> 32));
    if (d < depth[k]) {
        depth[k] = d;
        index_map[k] = (int32_t)id;
    }
}

__device__ __forceinline__ bool quad_nonzero3(const float* __restrict__ m, int wd, int x, int y) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (m[((size_t)y * wd + x) * 3 + c] == 0.f) return false;
        if (m[((size_t)y * wd + x + 1) * 3 + c] == 0.f) return false;
        if (m[((size_t)(y + 1) * wd + x) * 3 + c] == 0.f) return false;
        if (m[((size_t)(y + 1) * wd + x + 1) * 3 + c] == 0.f) return false;
    }
    return true;
}

__device__ __forceinline__ float clamp1(float v) { return v > 1.f ? 1.f : (v < -1.f ? -1.f : v); }

__global__ void __launch_bounds__(256)
k_fuse_surfels(int P, int ht, int wd, const float* __restrict__ intr, const float* __restrict__ view,
               const float* __restrict__ proj, const float* __restrict__ vmap, const float* __restrict__ nmap,
               const float* __restrict__ dmap, const uint8_t* __restrict__ fmask, const int32_t* __restrict__ imap,
               float* __restrict__ points, float* __restrict__ rotations, float* __restrict__ sigma2,
               uint8_t* __restrict__ inview, uint8_t* __restrict__ surface, float dist_thres, float alpha_p,
               float alpha_n) {
    __shared__ FuseCam c;
    load_fuse_cam(c, view, proj, intr, wd, ht);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    inview[i] = 0;
    surface[i] = 0;
    const float pw[3] = {points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]};
    const float4 q4 = *(reinterpret_cast<const float4*>(rotations) + i);
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float pc[3], coord[2], zax[3];
    Mat3 R;
    if (!fuse_common(c, pw, q, pc, coord, R, zax)) return;
    inview[i] = 1;
    {
        const int x = __float2int_rd(coord[0]), y = __float2int_rd(coord[1]);
        if (x < 0 || x + 1 >= wd || y < 0 || y + 1 >= ht) return;
        if (!quad_nonzero3(nmap, wd, x, y)) return;
        if (!fmask[(size_t)y * wd + x] || !fmask[(size_t)y * wd + x + 1] || !fmask[(size_t)(y + 1) * wd + x] ||
            !fmask[(size_t)(y + 1) * wd + x + 1])
            return;
    }
    const int rx = __float2int_rn(coord[0]), ry = __float2int_rn(coord[1]);
    const int ix = min(max(rx, 0), wd - 1), iy = min(max(ry, 0), ht - 1);
    const size_t nidx = (size_t)iy * wd + ix;
    float nw[3] = {R.m[0][2], R.m[1][2], R.m[2][2]};
    {
        const float inv = 1.0f / sqrtf(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
        nw[0] *= inv; nw[1] *= inv; nw[2] *= inv;
    }
    float nc[3] = {nmap[3 * nidx], nmap[3 * nidx + 1], nmap[3 * nidx + 2]};
    {
        const float inv = 1.0f / sqrtf(nc[0] * nc[0] + nc[1] * nc[1] + nc[2] * nc[2]);
        nc[0] *= inv; nc[1] *= inv; nc[2] *= inv;
    }
    const float vc[3] = {vmap[3 * nidx], vmap[3 * nidx + 1], vmap[3 * nidx + 2]};
    const float vd[3] = {pw[0] - vc[0], pw[1] - vc[1], pw[2] - vc[2]};
    const int sid = imap[(size_t)ry * wd + rx];
    if (sid >= 0 && sid == i) surface[i] = 1;
    if (sqrtf(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2]) > dist_thres) return;

    const float s2p = sigma2[2 * (size_t)i], s2n = sigma2[2 * (size_t)i + 1];
    const float d = dmap[nidx];
    const float s2pz = (alpha_p * d) * (alpha_p * d), s2nz = (alpha_n * d) * (alpha_n * d);
    const float s2p_new = 1.f / (1.f / s2pz + 1.f / s2p), s2n_new = 1.f / (1.f / s2nz + 1.f / s2n);
    const float lp = 1.f / s2pz, ln = 1.f / s2nz;
    float xn[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        points[3 * (size_t)i + k] = s2p_new * (pw[k] / s2p + lp * vc[k]);
        xn[k] = s2n_new * (nw[k] / s2n + ln * nc[k]);
    }
    sigma2[2 * (size_t)i] = s2p_new;

    const double angle = (double)(acosf(clamp1(nw[0] * nc[0] + nw[1] * nc[1] + nw[2] * nc[2])) * 180) / 3.1415926;
    if (!(angle < 60)) return;
    float inv = 1.0f / sqrtf(xn[0] * xn[0] + xn[1] * xn[1] + xn[2] * xn[2]);
    const float nn[3] = {xn[0] * inv, xn[1] * inv, xn[2] * inv};
    const float cr[3] = {nw[1] * nn[2] - nw[2] * nn[1], nw[2] * nn[0] - nw[0] * nn[2], nw[0] * nn[1] - nw[1] * nn[0]};
    inv = 1.0f / sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    const float n12[3] = {cr[0] * inv, cr[1] * inv, cr[2] * inv};
    const float ct = clamp1(nw[0] * nn[0] + nw[1] * nn[1] + nw[2] * nn[2]);
    const double theta = (double)(acosf(ct) * 180) / 3.1415926;
    if (theta < 1) return;
    const float ang = acosf(ct);
    const float nrm = sqrtf(n12[0] * n12[0] + n12[1] * n12[1] + n12[2] * n12[2]);
    const float ux = n12[0] / nrm, uy = n12[1] / nrm, uz = n12[2] / nrm;
    const float cs = cosf(ang), sn = sinf(ang);
    float R1[3][3];
    R1[0][0] = cs + ux * ux * (1 - cs);      R1[0][1] = ux * uy * (1 - cs) - uz * sn; R1[0][2] = ux * uz * (1 - cs) + uy * sn;
    R1[1][0] = uy * ux * (1 - cs) + uz * sn; R1[1][1] = cs + uy * uy * (1 - cs);      R1[1][2] = uy * uz * (1 - cs) - ux * sn;
    R1[2][0] = uz * ux * (1 - cs) - uy * sn; R1[2][1] = uz * uy * (1 - cs) + ux * sn; R1[2][2] = cs + uz * uz * (1 - cs);
    float R2[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) R2[a][b] = R.m[0][b] * R1[a][0] + R.m[1][b] * R1[a][1] + R.m[2][b] * R1[a][2];
    const float tr = R2[0][0] + R2[1][1] + R2[2][2];
    float qw, qx, qy, qz;
    if (tr > 0.0f) {
        const float ss = 0.5f / sqrtf(tr + 1.0f);
        qw = 0.25f / ss; qx = (R2[2][1] - R2[1][2]) * ss; qy = (R2[0][2] - R2[2][0]) * ss; qz = (R2[1][0] - R2[0][1]) * ss;
    } else if (R2[0][0] > R2[1][1] && R2[0][0] > R2[2][2]) {
        const float ss = 2.0f * sqrtf(1.0f + R2[0][0] - R2[1][1] - R2[2][2]);
        qw = (R2[2][1] - R2[1][2]) / ss; qx = 0.25f * ss; qy = (R2[0][1] + R2[1][0]) / ss; qz = (R2[0][2] + R2[2][0]) / ss;
    } else if (R2[1][1] > R2[2][2]) {
        const float ss = 2.0f * sqrtf(1.0f + R2[1][1] - R2[0][0] - R2[2][2]);
        qw = (R2[0][2] - R2[2][0]) / ss; qx = (R2[0][1] + R2[1][0]) / ss; qy = 0.25f * ss; qz = (R2[1][2] + R2[2][1]) / ss;
    } else {
        const float ss = 2.0f * sqrtf(1.0f + R2[2][2] - R2[0][0] - R2[1][1]);
        qw = (R2[1][0] - R2[0][1]) / ss; qx = (R2[0][2] + R2[2][0]) / ss; qy = (R2[1][2] + R2[2][1]) / ss; qz = 0.25f * ss;
    }
    *(reinterpret_cast<float4*>(rotations) + i) = make_float4(qw, qx, qy, qz);
    sigma2[2 * (size_t)i + 1] = s2n_new;
}

cudaError_t launch_project_surfels(int P, int ht, int wd, const float* points, const float* rotations,
                                   const uint8_t* stable, const float* intr, const float* view, const float* proj,
                                   unsigned long long* zbuf, int32_t* index_map, float* depth, cudaStream_t s) {
    const size_t n = (size_t)ht * wd;
    if (n == 0) return cudaSuccess;
    k_init_zbuffer<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, zbuf);
    if (P > 0) k_project_surfels<<<(P + 255) / 256, 256, 0, s>>>(P, ht, wd, points, rotations, stable, intr, view, proj, zbuf);
    k_unpack_zbuffer<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, zbuf, index_map, depth);
    return cudaGetLastError();
}

cudaError_t launch_fuse_surfels(int P, int ht, int wd, const float* intr, const float* view, const float* proj,
                                const float* vmap, const float* nmap, const float* dmap, const uint8_t* fmask,
                                const int32_t* imap, float* points, float* rotations, float* sigma2, uint8_t* inview,
                                uint8_t* surface, float dist_thres, float alpha_p, float alpha_n, cudaStream_t s) {
    if (P == 0) return cudaSuccess;
    k_fuse_surfels<<<(P + 255) / 256, 256, 0, s>>>(P, ht, wd, intr, view, proj, vmap, nmap, dmap, fmask, imap, points,
                                                   rotations, sigma2, inview, surface, dist_thres, alpha_p, alpha_n);
    return cudaGetLastError();
}
