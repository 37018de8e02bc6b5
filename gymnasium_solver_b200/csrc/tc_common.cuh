// tc_common.cuh — tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; no CUTLASS dependency).
//
// Operand convention used by the engine's tensor-core kernels: every fp32 matrix that feeds `tcgen05.mma.kind::tf32`
// lives in shared memory as 128-byte-swizzled SLABS of [rows][32 floats]:
//     byte(row, col) = slab(col / 32) * rows*128 + row*128 + (((col % 32) / 4) ^ (row & 7)) * 16 + (col % 4) * 4
// (slabs 1024-byte aligned).  That single physical layout is BOTH
//   * the canonical K-major SWIZZLE_128B operand with MN = row, K = col        (SBO = 1024 B between 8-row groups), and
//   * the canonical MN-major SWIZZLE_128B operand with MN = col, K = row       (LBO = slab stride between 32-col groups,
//                                                                               SBO = 1024 B between 8-row k-groups),
// so one copy of an activation tile serves the forward / dgrad GEMM (row = sample is M) and the wgrad GEMM (row = sample
// is the reduction dim K) without a transpose.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gs {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- swizzled slab addressing ------------------------------------------------------------------------------------
// float index of element (row, col) inside a tile of `rows` rows (cols in slabs of 32)
__device__ __forceinline__ int slab_index(int row, int col, int rows) {
    return (col >> 5) * rows * 32 + row * 32 + ((((col & 31) >> 2) ^ (row & 7)) << 2) + (col & 3);
}

// ---- descriptors ----------------------------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (sm_100 "version 1"): start addr >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// 1 << 46 | layout << 61.  layout: 0 none, 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
constexpr uint32_t kLayoutNone = 0, kLayoutSW128 = 2;

// 32-bit instruction descriptor for kind::tf32 with fp32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                      // c_format = F32
           | (2u << 7) | (2u << 10)       // a_format = b_format = TF32
           | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM management -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- single-thread issue --------------------------------------------------------------------------------------------------
// tcgen05.mma / commit take their operands from UNIFORM registers.  Issued under a thread-dependent branch (tid == 0) the
// compiler wraps every MMA in an ELECT / R2UR / BRA.U.ANY loop (~11 instructions and two uniform-datapath round trips
// each); under a warp-uniform branch + elect.sync, with operands derived from warp-uniform values, it is one UTCHMMA.
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- MMA -------------------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]  (A: 128 lanes x K columns of 32-bit elements); issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifndef GS_MBAR_HINT_NS
#define GS_MBAR_HINT_NS 0x400
#endif
// ---- mbarrier -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bounded wait: traps instead of hanging the GPU if the phase never completes (a kernel bug, never a legal state)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // suspend-time hint (ns): sleep, do not spin
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"((uint32_t)GS_MBAR_HINT_NS)
            : "memory");
        if (done) return;
#ifdef GS_MBAR_BACKOFF_NS
        __nanosleep(GS_MBAR_BACKOFF_NS);
#endif
    }
    __trap();
}

// ---- TMEM loads: 32 lanes x 32 bit, N consecutive columns per thread (thread i of warp w reads lane 32*(w%4)+i) -------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// TMEM store: thread i of warp w writes 16 consecutive columns of lane 32*(w%4)+i
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cp.async (global -> shared without a register round trip) ---------------------------------------------------------------
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    // .L2::64B: a 16-byte copy otherwise makes L2 fetch the whole 128-byte line from DRAM -- the 64-byte sample records of the update
    // kernel's gather cost 131 MB of DRAM reads per 1M-sample minibatch without it, 71.5 MB (= the bytes asked for) with it (ncu)
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// tf32 split: hi keeps the top 19 bits (what the tensor core reads), lo = x - hi (exact in fp32)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}
// round-to-nearest tf32 split: hi = rn_tf32(x) (cvt.rna), lo = x - hi is exact, signed and <= 2^-12 |x|; the tensor core
// then truncates lo to 11 bits, so a product hi*hi + lo*hi + hi*lo carries ~2^-23 relative error (fp32-like, unbiased)
__device__ __forceinline__ float tf32_rn(float x) {
    // round-half-away on the 13 dropped mantissa bits, two integer ops (cvt.rna.tf32.f32 costs four: it guards inf / nan,
    // which never reach an operand here)
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

}  // namespace tc
}  // namespace gs
