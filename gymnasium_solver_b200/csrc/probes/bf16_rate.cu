// bf16_rate.cu — SM cycles per tcgen05.mma.kind::f16 (bf16, K = 16) for the shapes of the bf16x3 update kernel, with 1, 2 and 4
// issuing warps (each with its own accumulator).  Operands are zeros: timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 bf16_rate.cu -o bf16_rate && ./bf16_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../tc_common.cuh"
using namespace gs::tc;
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) { return ((uint64_t)(0x40004040u) << 32) | (uint64_t)(((a >> 4) & 0x3FFFu) | 0x10000u); }
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__global__ void multi_kernel(int M, int n, int reps, int nw, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tmem = uniform(slot), s0 = smem_u32(sm), wu = uniform((uint32_t)warp);
    long long t0 = 0, t1 = 0, ti = 0;
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        t0 = clock64();
        if (wu < (uint32_t)nw) {
            if (elect_one()) {
                const uint32_t idesc = idesc_bf16(M, n);
                const uint32_t acc = tmem + wu * 128u;
                for (int i = 0; i < reps; ++i) {
                    const uint32_t koff = (uint32_t)(i & 7);
                    mma_f16(acc, desc_sw128(s0 + (koff >> 2) * 16384u + (koff & 3) * 32u), desc_sw128(s0 + 65536u + (koff >> 2) * 16384u + (koff & 3) * 32u), idesc, i > 0);
                }
                mma_commit(&bar[wu]);
            }
            ti = clock64();
            mbar_wait(&bar[wu], rep & 1);
        }
        t1 = clock64();
        __syncthreads();
        if (tid == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int shapes[][2] = {{128, 16}, {128, 64}, {128, 128}, {64, 16}, {64, 64}, {64, 80}};
    for (auto& s : shapes)
        for (int nw = 1; nw <= 4; nw *= 2) {
            const int reps = 48;
            multi_kernel<<<1, 128, 160 * 1024>>>(s[0], s[1], reps, nw, d);
            long long h[2];
            cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            printf("bf16 M=%3d N=%3d K=16, %d issuing warp(s) x %d MMAs: %6.1f cycles per MMA aggregate (%6.1f to issue)\n", s[0], s[1], nw, reps, (double)h[0] / (reps * nw), (double)h[1] / (reps * nw));
        }
    return 0;
}
