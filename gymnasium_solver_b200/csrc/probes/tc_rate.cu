// tc_rate.cu — how many SM cycles does one tcgen05.mma.kind::tf32 take for the shapes the update kernel issues?
// One CTA, one issuing thread; each test issues `reps` back-to-back MMAs into one accumulator (dependent chain, as in the
// kernel), commits, waits, and reports cycles per MMA.  Operands are whatever shared memory / TMEM holds (zeros): timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. tc_rate.cu -o tc_rate && ./tc_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../tc_common.cuh"

using namespace gs::tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    return ((uint64_t)(0x40004040u) << 32) | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | 0x10000u);
}

// kind 0: TS M=128 N=n ; kind 1: SS M=64 N=n ; kind 2: SS M=128 N=n.  `indep`: rotate over 4 accumulators instead of one.
__global__ void rate_kernel(int kind, int n, int reps, int indep, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(slot);
    const uint32_t s0 = smem_u32(sm);
    if (uniform((uint32_t)warp) == 0) {
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 3; ++rep) {     // last repetition is reported
            t0 = clock64();
            if (elect_one()) {
                const uint32_t idesc = make_idesc_tf32(kind == 1 ? 64 : 128, n, 0, 0);
                for (int i = 0; i < reps; ++i) {
                    const uint32_t acc = tmem + 128 + (indep ? (uint32_t)(i & 1) * 128u : 0u);
                    const uint32_t koff = (uint32_t)(i & 15);
                    const uint32_t boff = 65536u + (koff >> 2) * 16384u + (koff & 3) * 32u;
                    if (kind == 0) mma_tf32_ts(acc, tmem + (koff & 7) * 8, desc_sw128(s0 + boff), idesc, i > 0);
                    else mma_tf32(acc, desc_sw128(s0 + (koff >> 2) * 16384u + (koff & 3) * 32u), desc_sw128(s0 + boff), idesc, i > 0);
                }
                mma_commit(&bar);
            }
            const long long ti = clock64();
            mbar_wait(&bar, rep & 1);
            t1 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// `nw` warps issue `reps` MMAs each at the same time into their own accumulators: is the ~60-cycle floor per issuing warp or per SM?
__global__ void multi_kernel(int kind, int n, int reps, int nw, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(slot);
    const uint32_t s0 = smem_u32(sm);
    const uint32_t wu = uniform((uint32_t)warp);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        t0 = clock64();
        if (wu < (uint32_t)nw) {
            if (elect_one()) {
                const uint32_t idesc = make_idesc_tf32(kind == 1 ? 64 : 128, n, 0, 0);
                const uint32_t acc = tmem + 256 + wu * 64u;
                for (int i = 0; i < reps; ++i) {
                    const uint32_t koff = (uint32_t)(i & 15);
                    const uint32_t boff = 65536u + (koff >> 2) * 16384u + (koff & 3) * 32u;
                    if (kind == 0) mma_tf32_ts(acc, tmem + (wu & 1) * 64 + (koff & 7) * 8, desc_sw128(s0 + boff), idesc, i > 0);
                    else mma_tf32(acc, desc_sw128(s0 + (koff >> 2) * 16384u + (koff & 3) * 32u), desc_sw128(s0 + boff), idesc, i > 0);
                }
                mma_commit(&bar[wu]);
            }
            mbar_wait(&bar[wu], rep & 1);
        }
        t1 = clock64();
        __syncthreads();
        if (tid == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    struct { int kind, n; const char* name; } tests[] = {
        {0, 64, "TS  M=128 N=64  (fwd / dgrad)"}, {1, 72, "SS  M=64  N=72  (dW2|db2)"}, {1, 64, "SS  M=64  N=64"},
        {2, 16, "SS  M=128 N=16  (tail)"}, {1, 8, "SS  M=64  N=8"}, {2, 64, "SS  M=128 N=64"}, {0, 16, "TS  M=128 N=16"}, {0, 128, "TS  M=128 N=128"},
    };
    for (auto& t : tests)
        for (int indep = 0; indep < 2; ++indep) {
            const int reps = 48;
            rate_kernel<<<1, 128, 160 * 1024>>>(t.kind, t.n, reps, indep, d);
            long long h[2];
            cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("%s: %s\n", t.name, cudaGetErrorString(e)); return 1; }
            printf("%-34s %s: %6.1f cycles/MMA to completion, %6.1f cycles/MMA to issue (%d MMAs)\n", t.name, indep ? "2 accumulators" : "1 accumulator ",
                   (double)h[0] / reps, (double)h[1] / reps, reps);
        }
    cudaFuncSetAttribute(multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int kind = 0; kind < 2; ++kind)
        for (int nw = 1; nw <= 4; nw *= 2) {
            const int reps = 48;
            multi_kernel<<<1, 128, 160 * 1024>>>(kind, 64, reps, nw, d);
            long long h[2];
            cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("multi: %s\n", cudaGetErrorString(e)); return 1; }
            printf("%s N=64, %d warps issuing %d MMAs each: %6.1f cycles per MMA (aggregate)\n", kind ? "SS M=64 " : "TS M=128", nw, reps, (double)h[0] / (reps * nw));
        }
    return 0;
}
