// Microbenchmark: cycles per warp-wide LDS.128 / STS.128 as a function of how many distinct 16-byte addresses the 32 lanes
// touch and how they are arranged.  Answers: is a partially uniform 128-bit shared load served in 1 wavefront (general
// multicast) or in 4 quarter-warp passes?   nvcc -arch=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(int pattern, int iters, long long* out, float* sink) {
    extern __shared__ float4 s[];
    for (int i = threadIdx.x; i < 4096; i += 32) s[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncwarp();
    const int lane = threadIdx.x;
    int idx;
    switch (pattern) {
        case 0: idx = 0; break;                 // uniform
        case 1: idx = lane >> 4; break;         // 2 distinct, halves
        case 2: idx = lane & 1; break;          // 2 distinct, interleaved
        case 3: idx = lane >> 3; break;         // 4 distinct, quarter-uniform
        case 4: idx = lane & 3; break;          // 4 distinct, interleaved
        case 5: idx = lane >> 2; break;         // 8 distinct, groups of 4
        case 6: idx = lane & 7; break;          // 8 distinct, interleaved
        case 7: idx = lane; break;              // 32 distinct, consecutive (512 B)
        case 8: idx = (lane & 3) * 2; break;    // 4 distinct, 32 B apart (ktab pattern today)
        case 9: idx = lane * 33; break;         // 32 distinct, row stride 33 units (transposed park read)
        case 10: idx = (lane >> 2) * 36 + (lane & 3); break;  // today's pair-row read: 8 rows x 4 units
        default: idx = (lane & 15); break;      // 16 distinct interleaved
    }
    idx &= 1023;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s) + 16u * idx;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        float4 v0, v1, v2, v3;
        const uint32_t b = base + (uint32_t)(i & 3) * 16384u;   // same lane pattern, rotating 16 KB window
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(b) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(b) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v2.x), "=f"(v2.y), "=f"(v2.z), "=f"(v2.w) : "r"(b) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v3.x), "=f"(v3.y), "=f"(v3.z), "=f"(v3.w) : "r"(b) : "memory");
        a0.x += v0.x; a1.x += v1.y; a2.x += v2.z; a3.x += v3.w;
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 32 + lane] = a0.x + a1.x + a2.x + a3.x;
}
int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 8 * 1024); cudaMalloc(&sink, 4 * 32 * 1024);
    const int iters = 20000;
    const char* names[] = {"uniform", "2 halves", "2 interleaved", "4 quarter-uniform", "4 interleaved", "8 groups-of-4",
                           "8 interleaved", "32 consecutive", "4 x 32B apart", "32 stride-33", "8 rows x 4 (stride 36)", "16 interleaved"};
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int warps = 1; warps <= 3; warps += 2)
        for (int p = 0; p < 12; p++) {
            // `warps` one-warp CTAs per SM (148 SMs): cycles per LDS.128 per warp, and per SM
            k<<<148 * warps, 32, 65536>>>(p, iters, d, sink);
            cudaDeviceSynchronize();
            long long h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("warps/SM %2d  pattern %-24s  cycles per LDS.128 per warp %.2f  -> per SM %.2f\n", warps, names[p],
                   (double)h[0] / (4.0 * iters), (double)h[0] / (4.0 * iters) / warps);
        }
    return 0;
}
