// f16x3.cuh — what the fp16x3 tensor-core kernels (update_f16.cu, collect_f16.cu) share: the operand layout, descriptors, the
// hi / lo split, and the staging of an MLP's weights into operand tiles.
//
// Layout (every tile, activations and weights): [rows][64 * slabs] fp16 in 64-column slabs of 128-byte rows, 16-byte chunk c of row r
// stored at chunk c ^ (r & 7) -- the SWIZZLE_128B pattern, so the same bytes are a K-major operand (MN = row) and an MN-major operand
// (MN = column); probes/mma16_probe.cu holds the known-answer tests of every descriptor form used.
#pragma once

#include "mlp_tile.cuh"
#include "tc_common.cuh"

namespace gs {
namespace hfu {

using namespace tc;

constexpr int kRows = 128;               // rows per tile == TMEM lanes
constexpr uint32_t kSlab = 128 * 128;    // one [128][64] fp16 slab
constexpr int kMaxD = 7;                 // x16 = [x_hi (cols 0..) | x_lo (cols 7..) | 1 1 (cols 14, 15)]

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {   // a_format = b_format = 0 (F16), fp32 accumulation
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// SWIZZLE_128B descriptor, SBO = 1024 B (8-row groups); LBO = distance between 64-element slabs of an MN-major operand
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo = 16u) {
    return ((uint64_t)(0x40004040u) << 32) | (uint64_t)(((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16));
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory: rows = lanes, two fp16 per 32-bit column, even k in the low half (probes/ts16_probe.cu)
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// thread i of warp w writes 8 consecutive 32-bit columns of lane 32*(w%4)+i
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// packed fp16 pair -> its two values
__device__ __forceinline__ void unpack_pair(uint32_t p, float& a, float& b) {
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(a), "=f"(b) : "r"(p));
}
// (a, b) -> packed fp16 pairs: hi = {fp16(a) low half, fp16(b) high half} (saturating), lo = the rounded remainders
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    float ah, bh;
    unpack_pair(hi, ah, bh);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
}
__device__ __forceinline__ uint32_t f16_bits(float x) {   // fp16(x) in the low half
    uint32_t p;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(0.f), "f"(x));
    return p & 0xFFFFu;
}
__device__ __forceinline__ float f16_hi(float x) { float a, b; unpack_pair(f16_bits(x), a, b); return a; }
constexpr uint32_t kOnes2 = 0x3C003C00u;   // {1.0, 1.0}

// byte offset of fp16 element (r, c) in a [rows][64 * slabs] tile
__device__ __forceinline__ uint32_t tile_off(int r, int c, int rows) {
    return (uint32_t)(c >> 6) * (uint32_t)rows * 128u + (uint32_t)r * 128u + (uint32_t)(((((c & 63) >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
}


// An MLP's weights as operand tiles (once per launch):
//   w2hi / w2lo  [H][H]   W2[j][k] split into hi / lo: K-major B of the forward pass, MN-major B of dgrad
//   ws           [H][64]  four 16-column groups per output unit j:
//                0: [W1_hi(d) at d | W1_hi(d) at 7+d | b1_hi b1_lo at 14, 15]   x16 . ws0 = (x_hi + x_lo) . W1_hi + b1
//                1: [W1_lo(d) at d]                                              x16 . ws1 = x_hi . W1_lo
//                2: [b2_hi b2_lo at 14, 15]                                      x16 . ws2 = b2
//                3: [Wh_hi(r) at r | Wh_hi(r) at 4+r | Wh_lo(r) at 8+r]          head rows r < 4 (policy logits, then the value row)
template <int H>
__device__ __forceinline__ void stage_w2(const MlpDev& m, unsigned char* w2hi, unsigned char* w2lo, int tid, int n_threads) {
#pragma unroll 4
    for (int i = tid; i < H * H / 4; i += n_threads) {      // independent 16-byte loads, several in flight per thread
        const int j = i / (H / 4), k = 4 * (i % (H / 4));
        const float4 w = __ldg(reinterpret_cast<const float4*>(m.w2 + j * H) + (i % (H / 4)));
        uint32_t h0, l0, h1, l1;
        split_pair(w.x, w.y, h0, l0);
        split_pair(w.z, w.w, h1, l1);
        const uint32_t o = tile_off(j, k, H);
        *reinterpret_cast<uint2*>(w2hi + o) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(w2lo + o) = make_uint2(l0, l1);
    }
}
template <int H>
__device__ __forceinline__ void stage_ws(const MlpDev& m, unsigned char* ws, int tid) {
    const int D = m.D, A = m.A;
    if (tid < H) {                                            // WS row j: every load of the row is independent -> one round trip
        const int j = tid;
        float w[kMaxD], whv[4];
#pragma unroll
        for (int d = 0; d < kMaxD; ++d) w[d] = d < D ? __ldg(m.w1 + j * D + d) : 0.f;
        const float b1v = __ldg(m.b1 + j), b2v = __ldg(m.b2 + j);
#pragma unroll
        for (int r = 0; r < 4; ++r) whv[r] = r < A ? __ldg(m.wp + r * H + j) : ((r == A && m.has_value) ? __ldg(m.wv + j) : 0.f);
        uint32_t e[64];
#pragma unroll
        for (int q = 0; q < 64; ++q) e[q] = 0u;
#pragma unroll
        for (int d = 0; d < kMaxD; ++d) {
            const uint32_t hi = f16_bits(w[d]);
            e[d] = hi; e[7 + d] = hi; e[16 + d] = f16_bits(w[d] - f16_hi(w[d]));
        }
        e[14] = f16_bits(b1v); e[15] = f16_bits(b1v - f16_hi(b1v));
        e[32 + 14] = f16_bits(b2v); e[32 + 15] = f16_bits(b2v - f16_hi(b2v));
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t hi = f16_bits(whv[r]);
            e[48 + r] = hi; e[52 + r] = hi; e[56 + r] = f16_bits(whv[r] - f16_hi(whv[r]));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(ws + j * 128 + ((c ^ (j & 7)) << 4)) =
                make_uint4(e[8 * c] | (e[8 * c + 1] << 16), e[8 * c + 2] | (e[8 * c + 3] << 16), e[8 * c + 4] | (e[8 * c + 5] << 16), e[8 * c + 6] | (e[8 * c + 7] << 16));
    }
}
template <int H>
__device__ __forceinline__ void stage_weights(const MlpDev& m, unsigned char* w2hi, unsigned char* w2lo, unsigned char* ws, int tid, int n_threads) {
    stage_w2<H>(m, w2hi, w2lo, tid, n_threads);
    stage_ws<H>(m, ws, tid);
}

}  // namespace hfu
}  // namespace gs
