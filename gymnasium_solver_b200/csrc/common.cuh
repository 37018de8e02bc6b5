// common.cuh — shared device/host helpers of the B200 rollout engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gs_engine.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gs_engine kernels are written for sm_100a (B200) only"
#endif

namespace gs {

// ---- error plumbing (thread-local message, negative return codes) -----------------------------------
void set_error(const char* fmt, ...);
#define GS_FAIL(...)                \
    do {                            \
        ::gs::set_error(__VA_ARGS__); \
        return -1;                  \
    } while (0)
#define GS_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) GS_FAIL("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define GS_LAUNCH_CHECK()                                                                   \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) GS_FAIL("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

int sm_count(int device);

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw — SC'11), counter-based: no RNG state in HBM ----------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

constexpr uint32_t kTagReset = 0x5E5E0000u;   // env reset noise stream
constexpr uint32_t kTagAction = 0xAC700000u;  // action sampling stream

// 53-bit uniform in [0,1) from two words (the construction numpy's random() uses)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
// 24-bit uniform in [0,1)
__device__ __forceinline__ float u24(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float action_uniform(uint64_t seed, uint64_t gid, uint64_t vec_step) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)vec_step,
                                             kTagAction | ((uint32_t)(vec_step >> 32) & 0xFFFFFu)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return u24(r.x);
}

// ---- IEEE fp64/fp32 arithmetic that ptxas may not contract into FMAs (bit parity with numpy / gcc -ffp-contract=off)
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// ---- reductions -----------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// block-wide sum of doubles; result valid in thread 0.  scratch: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < nw ? scratch[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// ---- streaming loads / stores (data touched once: keep it out of L1) ---------------------------------
__device__ __forceinline__ float ldg_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ uint8_t ldg_stream(const uint8_t* p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(float* p, float v) { __stcs(p, v); }

// ---- pseudo-random bijection on [0, len): 4-round alternating Feistel network on exactly ceil(log2(len)) bits + cycle walking.
// Replaces the n_epochs*N*T Python index list of MultiPassRandomSampler (utils/samplers.py:29-34) with O(1) per sample.
// The halves may differ by one bit (L: bits/2 high bits, R: the rest); round r XORs one half with a keyed hash of the other
// (even rounds L ^= F(R), odd rounds R ^= F(L)), so every round is invertible whatever the widths.  A power-of-two length
// (the usual rollout: N*T = 2^k) needs no cycle walking at all; otherwise fewer than half of the draws walk once more.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// domain width of the permutation of [0, len): the smallest bits >= 2 with 2^bits >= len
__host__ __device__ __forceinline__ int feistel_bits(uint64_t len) {
    int bits = 2;
    while ((1ull << bits) < len) bits += 1;
    return bits;
}
// `bits` = feistel_bits(len) when the caller has it (a per-launch constant: the search costs ~60 instructions per sample), else 0
__host__ __device__ __forceinline__ uint64_t feistel_permute(uint64_t x, uint64_t len, uint64_t key, int bits = 0) {
    if (bits <= 0) bits = feistel_bits(len);
    const int lb = bits >> 1, rb = bits - lb;
    const uint32_t lmask = (uint32_t)((1ull << lb) - 1), rmask = (uint32_t)((1ull << rb) - 1);
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    do {
        uint32_t L = (uint32_t)(x >> rb) & lmask, R = (uint32_t)x & rmask;
        L ^= mix32(R * 0x9E3779B1u + k0) & lmask;
        R ^= mix32(L * 0x85EBCA6Bu + k1) & rmask;
        L ^= mix32(R * 0xC2B2AE35u + (k0 ^ 0x27D4EB2Fu)) & lmask;
        R ^= mix32(L * 0x165667B1u + (k1 ^ 0x9E3779B9u)) & rmask;
        x = ((uint64_t)L << rb) | R;
    } while (x >= len);
    return x;
}

}  // namespace gs
