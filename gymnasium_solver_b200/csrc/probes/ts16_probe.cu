// ts16_probe.cu — known-answer test of tcgen05.mma.kind::f16 with the A operand in TENSOR MEMORY (rows = lanes, two fp16 per 32-bit column).
// D[128][128] = A[128][64] . B[128][64]^T with A written by tcgen05.st (one row per thread), B a K-major SWIZZLE_128B tile in shared
// memory; four K = 16 steps: A at column base + 8 kk, B at +32 kk bytes.  Prints the packing that matches (low half = even k, or odd k).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I.. -o ts16_probe ts16_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../f16x3.cuh"

using namespace gs::tc;
using namespace gs::hfu;

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void probe(const uint32_t* __restrict__ a_packed /* [128][32] */, const unsigned char* __restrict__ b_img /* 16 KB */, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint32_t tmem_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 16384 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_s, T = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 32; c += 16) {
        float v[16];
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a_packed[tid * 32 + c + i]);
        tmem_st16(T + 128 + c, v);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        for (int kk = 0; kk < 4; ++kk) mma_f16_ts(tmem, tmem + 128 + 8 * kk, desc(smem_u32(sm) + 32u * kk), idesc_f16(128, 128, 0, 0), kk ? 1u : 0u);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c = 0; c < 128; c += 16) {
        float v[16];
        tmem_ld16(T + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[tid * 128 + c + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static uint16_t hb(float x) { __half h = __float2half(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float hr(float x) { return __half2float(__float2half(x)); }
static size_t tile_byte(int r, int c, int rows) { return (size_t)(c / 64) * rows * 128 + (size_t)r * 128 + (size_t)((((c % 64) / 8) ^ (r & 7)) * 16) + (c % 8) * 2; }

int main() {
    std::vector<float> A(128 * 64), B(128 * 64);
    srand(3);
    for (auto& x : A) x = hr(((rand() % 2001) - 1000) / 1000.0f);
    for (auto& x : B) x = hr(((rand() % 2001) - 1000) / 1000.0f);
    std::vector<unsigned char> bimg(16384, 0);
    for (int n = 0; n < 128; ++n) for (int k = 0; k < 64; ++k) { uint16_t u = hb(B[n * 64 + k]); memcpy(&bimg[tile_byte(n, k, 128)], &u, 2); }
    std::vector<double> ref(128 * 128, 0.0);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)A[m * 64 + k] * B[n * 64 + k]; ref[m * 128 + n] = s; }
    uint32_t* d_a; unsigned char* d_b; float* d_o;
    cudaMalloc(&d_a, 128 * 32 * 4); cudaMalloc(&d_b, 16384); cudaMalloc(&d_o, 128 * 128 * 4);
    cudaMemcpy(d_b, bimg.data(), 16384, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    int fails = 0;
    for (int packing = 0; packing < 2; ++packing) {
        std::vector<uint32_t> ap(128 * 32);
        for (int m = 0; m < 128; ++m) for (int c = 0; c < 32; ++c) {
            const uint16_t e = hb(A[m * 64 + 2 * c]), o = hb(A[m * 64 + 2 * c + 1]);
            ap[m * 32 + c] = packing == 0 ? ((uint32_t)o << 16 | e) : ((uint32_t)e << 16 | o);
        }
        cudaMemcpy(d_a, ap.data(), ap.size() * 4, cudaMemcpyHostToDevice);
        probe<<<1, 128, 16384>>>(d_a, d_b, d_o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("packing %d: CUDA error %s\n", packing, cudaGetErrorString(e)); return 2; }
        std::vector<float> o(128 * 128);
        cudaMemcpy(o.data(), d_o, o.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < 128 * 128; ++i) worst = fmax(worst, fabs(o[i] - ref[i]));
        printf("A in TMEM, packing %d (low half = %s k): max|err| %.3e -> %s\n", packing, packing == 0 ? "even" : "odd", worst, worst < 1e-4 ? "MATCH" : "no");
        if (packing == 0 && worst >= 1e-4) ++fails;
    }
    printf("%s\n", fails ? "TS16 PROBE: even-k-low packing does NOT match" : "TS16 PROBE OK (even k in the low half, column = k / 2, lane = row)");
    return 0;
}
