// update_tc.cu — tcgen05 / TMEM version of the fused minibatch update for the 64x64 MLP (PPO and REINFORCE).
//
// Same contract as update_kernels.cu::update_kernel (gather -> forward -> loss + metrics -> backward -> per-CTA partial
// gradients), but the 64x64 GEMMs and every cross-sample reduction run on the 5th-gen tensor cores.  Four MMA groups per
// 128-sample tile (3xTF32: hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM):
//
//   fwd(u)    z2[s][j]   = sum_k h1[s][k]  W2[j][k]      M=128 (samples) N=64 K=64    A: TMEM (h1 rows),  B: W2   smem
//   dgrad(t)  dh1[s][k]  = sum_j dz2[s][j] W2[j][k]      M=128           N=64 K=64    A: TMEM (dz2 rows), B: W2^T smem
//   dW2(t)    dW2[j][k] += sum_s dz2[s][j] h1[s][k]      M=64  N=72 K=128 (samples)   A: dz2^T (P), B: h1^T + a ones row (S):
//             db2[j]    += sum_s dz2[s][j]                                             column 64 of the same accumulator
//   tail(t,u) dW1[j][d] += sum_s dz1(t)[s][j] x(t)[s][d], db1 (ones row of x^T)        M=128 N=16 K=128
//             dWh[r][k] += sum_s g(u)[s][r] h2(u)[s][k]                                A: [dz1^T(t) ; h2^T(u)] (P;S), B: [x^T(t) ; g^T(u)]
//             (one MMA group, block-diagonal use of a 128x16 accumulator; the off-diagonal blocks are never read)
//
// Layer 1, heads, softmax, loss and the activation derivatives are row-local fp32 SIMT work on registers.  Shared-memory
// operands use the K-major SWIZZLE_128B slab layout of tc_common.cuh; the transposed operands ([feature][sample]) are
// written straight from registers, conflict-free (a warp writes 128 contiguous bytes of one row).  P and S are the two
// 64-row halves of ONE 128-row tile (+ 8 rows holding the ones row behind S), so the same bytes serve as M=64 / N=72
// operands of dW2 and as the M=128 operand of the tail group.
//
// Schedule: 1 CTA per SM = 16 compute warps (thread = sample row x 16-column chunk) + 1 MMA-issuing warp, 227 KB shared
// memory, all 512 TMEM columns.  tcgen05.mma issue blocks the issuing warp for as long as the tensor pipe is busy (the
// queue is shallow: measured with the clock64 trace), so it gets a warp of its own; the compute warps synchronise among
// themselves on a named barrier and hand operands over through mbarriers.  The loop is software-pipelined over tiles:
// iteration `it` runs the BACKWARD phase of tile t and the FORWARD phase of the next tile u in one instruction stream, so
// every MMA group has independent SIMT work behind it:
//
//   B1(t)  z2 -> dz2 -> TMEM A[cur], layer 1 again -> h1 ; wait tail(t-1) ; dz2^T -> P, h1^T -> S, x^T(t) -> Y
//          owners: sample(u) -> TMEM scratch
//   ---- sync 1 ---- ready[0]:  MMA warps issue dgrad(t) [A[cur] -> acc[cur]] and dW2(t) [P,S]
//   F1(u)  layer 1 -> h1(u) -> TMEM A[nxt]
//   ---- sync 2 ---- ready[1]:  MMA warp issues fwd(u) [A[nxt] -> acc[nxt]]
//          wait dgrad(t), dW2(t) ; dz1^T(t) -> P
//   F2(u)  wait fwd(u); z2 -> h2 -> partial head outputs -> TMEM scratch
//   ---- sync 3 ----
//   F3(u)  loss, d(loss)/d(heads) -> g (every chunk thread of the row; metrics by the chunk-1 thread), g^T(u) -> Y, h2^T(u) -> S
//   ---- sync 4 ---- ready[2]:  MMA warp issues tail(t,u) [P;S, Y]
//
// TMEM A operands and the z2 / dh1 accumulator are double-buffered (cur/nxt); the transposed operand tile is single
// (shared memory is the limit).  Per-row exchanges between the four chunk threads of a row (head partial sums, gathered
// sample) go through spare TMEM columns: all four warps of a quadrant address the same lanes.  All MMAs are issued by one
// thread in a fixed order, so every accumulation order is fixed (deterministic).
#include <type_traits>

#include "mlp_tile.cuh"
#include "tc_common.cuh"
#include "update_shared.cuh"

namespace gs {

using namespace tc;

namespace tcu {
constexpr int kRows = 128;              // samples per tile == TMEM lanes
constexpr int kCompute = 512;           // compute threads: thread = (row, 16-column chunk); warp w -> rows 32*(w&3).., chunk w>>2
constexpr int kT = kCompute + 64;       // + two MMA-issuing warps (16: dgrad, fwd; 17: dW2, tail)
// shared memory map (bytes).  Transposed operand tile: 4 slabs (32 samples each) x rows x 128 B.
constexpr int kHiRows = 136;            // P rows 0..63, S rows 64..127, ones row 128 (+7 zero rows: N is a multiple of 8)
constexpr int kHiSlab = kHiRows * 128;  // 17408 B
constexpr int kLoSlab = 128 * 128;      // 16384 B (no ones rows: their lo part is 0, that pass runs with N=64)
constexpr int oTH = 0;                  // hi tile  [4][136][32] fp32
constexpr int oTL = 4 * kHiSlab;        // lo tile  [4][128][32] fp32
constexpr int oW2 = oTL + 4 * kLoSlab;  // W2  hi 16K, lo 16K      [64 j][64 k]
constexpr int oW2T = oW2 + 32768;       // W2^T                    [64 k][64 j]
constexpr int oY = oW2T + 32768;        // [x^T ; g^T]: hi 8K, lo 8K   [16 rows][128 samples]; row 7 = ones, rows 8.. = g^T
constexpr int oMisc = oY + 16384;
constexpr int oW1 = oMisc;              // [64][8] fp32
constexpr int oB1 = oW1 + 2048;
constexpr int oB2 = oB1 + 256;
constexpr int oWH = oB2 + 256;          // [4][64]
constexpr int oBH = oWH + 1024;         // [4]
constexpr int oBar = oBH + 16;          // 7 mbarriers
constexpr int oTmem = oBar + 64;
constexpr int oRed = oTmem + 16;        // PM_N doubles + 16 floats (block reductions) + 8 floats (normalisation constants, 1/B)
constexpr int oStage = oRed + 320;      // gather staging [128 rows][16] fp32 (cp.async destination)
static_assert(PM_N * 8 + 16 * 4 + 8 * 4 <= 320, "reduction scratch");
constexpr int oScal1 = oStage + 128 * 64;   // second buffer of the 5 sample scalars [128 rows][5] (odd tiles; even tiles use slots 8..12 of the row)
constexpr int oOffN = oScal1 + 128 * 20;    // translated offset of each row's sample in the tile after next [128] u32
constexpr int kSmemBytes = oOffN + 128 * 4;
static_assert(oTL % 1024 == 0 && oW2 % 1024 == 0 && oY % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(kSmemBytes <= 232448, "shared memory budget");
// TMEM columns: A[b] = 128*b (hi +0, lo +64); acc[b] = 256 + 64*b; dW2|db2 (72 used of 80); tail accumulator; per-row scratch
constexpr uint32_t cA = 0, cAcc = 256, cW2 = 384, cC = 464, cOP = 480, cC2 = 496, kTmemCols = 512;   // cC2: hi*lo pass of the tail (second MMA warp)
enum { BAR_FWD = 0, BAR_D, BAR_W, BAR_T, RDY_0, RDY_1, RDY_2 };   // MMA-complete barriers; operands-ready barriers
// barrier among the 512 compute threads only (the MMA warp never joins it)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tcu

// Development aid (build with GS_NVCC_EXTRA=-DGS_TC_TRACE): per-warp clock64() stamps of one pipeline iteration of one CTA,
// read back with gs_debug_tc_trace().  Not part of the ABI; absent from normal builds.
#ifdef GS_TC_TRACE
__device__ long long g_tc_trace[18][24];
#define GS_TR(k) do { if (blockIdx.x == 7 && it == 5 && lane == 0) g_tc_trace[warp][k] = clock64(); } while (0)
#else
#define GS_TR(k) do { } while (0)
#endif

// operands are stored as (hi, lo) = (rn_tf32(x), x - hi): see tc_common.cuh::tf32_rn

// 16 values of the thread's row chunk -> TMEM A columns (hi, lo)
__device__ __forceinline__ void chunk_to_tmem(uint32_t a_addr /* lane | first column of the A buffer */, int c, const float (&v)[16]) {
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { h[i] = tf32_rn(v[i]); l[i] = v[i] - h[i]; }
    tmem_st16(a_addr + 16 * c, h);
    tmem_st16(a_addr + 64 + 16 * c, l);
}
// the chunk -> column s of a 64-row region of the transposed tile (rows 16c .. 16c+15), hi and lo copies.
// A warp writes 128 contiguous bytes per row: conflict-free.
__device__ __forceinline__ void chunk_to_transposed(float* hi, float* lo, int base_hi, int base_lo, const int (&xo)[8], int c, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int o = (16 * c + i) * 32 + xo[i & 7];
        const float h = tf32_rn(v[i]);
        hi[base_hi + o] = h;
        lo[base_lo + o] = v[i] - h;
    }
}

// pre-activations of layer-1 neurons 16c .. 16c+15 for one observation
template <bool D4>
__device__ __forceinline__ void layer1_chunk(const float* w1s, const float* b1s, const float (&x)[8], int c, float (&z)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = 16 * c + i;
        const float4 wa = *reinterpret_cast<const float4*>(w1s + j * 8);
        float acc = b1s[j];
        acc = fmaf(wa.x, x[0], acc); acc = fmaf(wa.y, x[1], acc); acc = fmaf(wa.z, x[2], acc); acc = fmaf(wa.w, x[3], acc);
        if (!D4) {
            const float4 wb = *reinterpret_cast<const float4*>(w1s + j * 8 + 4);
            acc = fmaf(wb.x, x[4], acc); acc = fmaf(wb.y, x[5], acc); acc = fmaf(wb.z, x[6], acc); acc = fmaf(wb.w, x[7], acc);
        }
        z[i] = acc;
    }
}

// activation statistics of one chunk (utils/models.py:121-146): sum, sum of squares, per-neuron |z| < 1e-6 counts.
// The dead test is one FMNMX per element; the counting path runs only when some row of the warp trips it, and counts with
// warp ballots into a lane-distributed register (lane i: neuron 16c + i) -- no atomics in the loop: a unit that is dead for
// every sample would otherwise serialise one global atomic per sample on a single address (measured: 50x slower kernel).
__device__ __forceinline__ void chunk_stats(const float (&z)[16], bool valid, int lane, float& s, float& q, uint32_t& dead_cnt) {
    float mn = 1.0f;
    if (valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { s += z[i]; q = fmaf(z[i], z[i], q); mn = fminf(mn, fabsf(z[i])); }
    }
    if (__any_sync(0xffffffffu, mn < 1e-6f)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t hits = __popc(__ballot_sync(0xffffffffu, valid && fabsf(z[i]) < 1e-6f));
            if (lane == i) dead_cnt += hits;
        }
    }
}

// K-major SWIZZLE_128B descriptor with SBO = 1024 B, LBO = 16 B: only the 14-bit start-address field changes between MMAs
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    return ((uint64_t)(0x40004040u) << 32) | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | 0x10000u);
}
// byte offset of k-step kk (8 samples) inside a transposed tile whose 32-sample slabs are `slab` bytes apart
__device__ __forceinline__ uint32_t kstep(int kk, uint32_t slab) { return (uint32_t)(kk >> 2) * slab + (uint32_t)(kk & 3) * 32u; }

// 3xTF32 groups, each issued by ONE thread ------------------------------------------------------------------------------------
// A from TMEM (M=128 rows = lanes), B K-major smem [64 rows][64 k] (lo copy 16 KB after hi)
__device__ __forceinline__ void issue_ts_64x64(uint32_t tmem_base, uint32_t a_col, uint32_t b_hi_addr, uint32_t acc_col) {
    const uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
    uint32_t accumulate = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t acol = tmem_base + a_col + (pass == 1 ? 64u : 0u);
        const uint32_t b0 = b_hi_addr + (pass == 2 ? 16384u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            mma_tf32_ts(tmem_base + acc_col, acol + kk * 8, desc_sw128(b0 + kstep(kk, 8192u)), idesc, accumulate);
            accumulate = 1;
        }
    }
}
// dW2 | db2: A = P (dz2^T, 64 rows), B = S + ones rows (h1^T, 72 rows; the lo pass has no ones rows: N = 64)
__device__ __forceinline__ void issue_dw2(uint32_t tmem_acc, uint32_t th, uint32_t tl, uint32_t accumulate) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t idesc = make_idesc_tf32(64, pass == 2 ? 64 : 72, 0, 0);
        const uint32_t a0 = pass == 1 ? tl : th, as = pass == 1 ? tcu::kLoSlab : tcu::kHiSlab;
        const uint32_t b0 = (pass == 2 ? tl : th) + 64u * 128u, bs = pass == 2 ? tcu::kLoSlab : tcu::kHiSlab;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            mma_tf32(tmem_acc, desc_sw128(a0 + kstep(kk, as)), desc_sw128(b0 + kstep(kk, bs)), idesc, accumulate);
            accumulate = 1;
        }
    }
}
// tail: A = [P ; S] (128 rows), B = Y (16 rows; lo copy 8 KB after hi)
// The three 3xTF32 passes are split over the two MMA warps (two accumulators, summed when they are flushed): passes
// [p0, p1) go to tmem_acc.  hi*hi + lo*hi -> one warp, hi*lo -> the other: the group completes in ~2/3 of the time.
__device__ __forceinline__ void issue_tail(uint32_t tmem_acc, uint32_t th, uint32_t tl, uint32_t y, uint32_t accumulate, int p0, int p1) {
    const uint32_t idesc = make_idesc_tf32(128, 16, 0, 0);
#pragma unroll
    for (int pass = p0; pass < p1; ++pass) {
        const uint32_t a0 = pass == 1 ? tl : th, as = pass == 1 ? tcu::kLoSlab : tcu::kHiSlab;
        const uint32_t b0 = y + (pass == 2 ? 8192u : 0u);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            mma_tf32(tmem_acc, desc_sw128(a0 + kstep(kk, as)), desc_sw128(b0 + kstep(kk, 2048u)), idesc, accumulate);
            accumulate = 1;
        }
    }
}

// Move the wgrad accumulators from TMEM into this CTA's fp32 partial-gradient vector (global, L2-resident) and let the next
// MMA groups restart from zero.  The tensor core accumulates with truncation, so a TMEM chain is kept to kFlushTiles tiles
// (bias ~1e-5 relative, measured); across chains the sums are round-to-nearest fp32 adds in a fixed order (deterministic).
//   dW2|db2 (M=64): row m lives in lane (m/16)*32 + m%16, i.e. lanes 0..15 of the warps of quadrant m/16.
//   tail (M=128):   lane = row; rows 0..63: cols 0..6 dW1[row][d], col 7 db1[row]; rows 64..127: cols 8+r dWh[r][row-64].
constexpr int kFlushTiles = 8;
__device__ __forceinline__ void flush_wgrad(uint32_t lane_addr, int quad, int chunk, int lane, int row, int D, int A, int has_value,
                                            float* __restrict__ out, bool first) {
    const ParamOffsets po = param_offsets(D, 64, 64, A, has_value);
    const int mrow = quad * 16 + (lane & 15);
    const bool owner = lane < 16;
    float4* dst = reinterpret_cast<float4*>(out + po.w2 + mrow * 64 + 16 * chunk);
    float4 old[4];
    const bool vec = ((po.w2 & 3) == 0);   // po.w2 = 64*D + 64 and the partial vector itself is 16 B aligned
    if (owner && !first) {
        if (vec) {
#pragma unroll
            for (int q = 0; q < 4; ++q) old[q] = dst[q];
        }
    }
    float v[16];
    tmem_ld16(lane_addr + tcu::cW2 + 16 * chunk, v);
    float s16[16];
    if (chunk < 2) tmem_ld16(lane_addr + (chunk == 0 ? tcu::cC : tcu::cW2 + 64), s16);   // chunk 0: tail block; chunk 1: db2 in col 0
    if (chunk == 0) {
        float t16[16];
        tmem_ld16(lane_addr + tcu::cC2, t16);                      // the tail's hi*lo pass has its own accumulator
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) s16[i] += t16[i];
    }
    tmem_ld_wait();
    auto put = [&](int64_t idx, float val) { out[idx] = first ? val : out[idx] + val; };
    if (owner) {
        if (vec) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 o = first ? make_float4(0.f, 0.f, 0.f, 0.f) : old[q];
                o.x += v[4 * q]; o.y += v[4 * q + 1]; o.z += v[4 * q + 2]; o.w += v[4 * q + 3];
                dst[q] = o;
            }
        } else {
            float* d1 = out + po.w2 + mrow * 64 + 16 * chunk;
#pragma unroll
            for (int i = 0; i < 16; ++i) d1[i] = first ? v[i] : d1[i] + v[i];
        }
        if (chunk == 1) put(po.b2 + mrow, s16[0]);
    }
    if (chunk == 0) {
        if (row < 64) {
#pragma unroll
            for (int d = 0; d < 7; ++d) if (d < D) put(po.w1 + row * D + d, s16[d]);
            put(po.b1 + row, s16[7]);
        } else {
            const int k = row - 64;
#pragma unroll
            for (int r = 0; r < 3; ++r) if (r < A) put(po.wp + r * 64 + k, s16[8 + r]);
            if (has_value) put(po.wv + k, A == 2 ? s16[10] : s16[11]);
        }
    }
}

template <int ALGO, bool TRACK, bool D4, int ACT>
__global__ void __launch_bounds__(tcu::kT, 1)
update_tc_kernel(MlpDev m, BatchDev b, HpDev hp, const double* __restrict__ adv_mom, const double* __restrict__ ret_mom,
                 const uint32_t* __restrict__ offs /* time-major offset of every minibatch position */, float* __restrict__ grad_partials,
                 int64_t pstride, double* __restrict__ metric_partials, uint32_t* __restrict__ dead) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* TH = reinterpret_cast<float*>(smraw + tcu::oTH);
    float* TL = reinterpret_cast<float*>(smraw + tcu::oTL);
    float* Yhi = reinterpret_cast<float*>(smraw + tcu::oY);
    float* Ylo = Yhi + 2048;
    float* w1s = reinterpret_cast<float*>(smraw + tcu::oW1);
    float* b1s = reinterpret_cast<float*>(smraw + tcu::oB1);
    float* b2s = reinterpret_cast<float*>(smraw + tcu::oB2);
    float* whs = reinterpret_cast<float*>(smraw + tcu::oWH);
    float* bhs = reinterpret_cast<float*>(smraw + tcu::oBH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + tcu::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smraw + tcu::oTmem);
    double* red = reinterpret_cast<double*>(smraw + tcu::oRed);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127;              // sample row of the tile == TMEM lane
    const int chunk = tid >> 7;             // 16-column chunk of the row this thread owns
    const int quad = warp & 3;              // TMEM lane quadrant of the warp
    const bool row_owner = chunk == 0;      // does the per-row gather and writes x^T / g^T

    const int A = m.A;

    // ---- one-time staging ------------------------------------------------------------------------------------------------
    if (warp == 0) tmem_alloc(tmem_slot, tcu::kTmemCols);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) mbar_init(&bars[i], i == tcu::BAR_T ? 2 : 1);   // the tail group is issued by both MMA warps
        fence_mbar_init();
    }
    {
        float* W2hi = reinterpret_cast<float*>(smraw + tcu::oW2);
        float* W2lo = W2hi + 4096;
        float* WThi = reinterpret_cast<float*>(smraw + tcu::oW2T);
        float* WTlo = WThi + 4096;
        for (int i = tid; i < 4096; i += tcu::kT) {
            const int j = i >> 6, k = i & 63;
            const float w = __ldg(m.w2 + i);
            const float wh = tf32_rn(w), l = w - wh;
            const int a = slab_index(j, k, 64), t = slab_index(k, j, 64);
            W2hi[a] = wh; W2lo[a] = l; WThi[t] = wh; WTlo[t] = l;
        }
        for (int i = tid; i < 512; i += tcu::kT) { const int j = i >> 3, d = i & 7; w1s[i] = d < m.D ? __ldg(m.w1 + j * m.D + d) : 0.f; }
        for (int i = tid; i < 64; i += tcu::kT) { b1s[i] = __ldg(m.b1 + i); b2s[i] = __ldg(m.b2 + i); }
        for (int i = tid; i < 256; i += tcu::kT) {
            const int r = i >> 6, k = i & 63;
            whs[i] = r < A ? __ldg(m.wp + r * 64 + k) : ((r == A && m.has_value) ? __ldg(m.wv + k) : 0.f);
        }
        if (tid < 4) bhs[tid] = tid < A ? __ldg(m.bp + tid) : ((tid == A && m.has_value) ? __ldg(m.bv) : 0.f);
        // the transposed tile and Y start as zeros: the first tail group (heads of the first tile only) reads all of them
        float4* z4 = reinterpret_cast<float4*>(smraw + tcu::oTH);
        for (int i = tid; i < (tcu::oW2 - tcu::oTH) / 16; i += tcu::kT) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < 4096; i += tcu::kT) Yhi[i] = 0.f;
        if (tid < PM_N + 8) red[tid] = 0.0;
    }
    __syncthreads();
    // transposed-store constants of this thread (sample column s = row): element (feature f, sample s) of a region sits at
    // base + f*32 + xo[f & 7] floats, base = slab(s) + (s & 3)
    const int chunk_s = (row & 31) >> 2;
    int xo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) xo[c] = (chunk_s ^ c) << 2;
    const int base_hi = (row >> 5) * (tcu::kHiSlab / 4) + (row & 3);
    const int base_lo = (row >> 5) * (tcu::kLoSlab / 4) + (row & 3);
    const int base_y = (row >> 5) * 512 + (row & 3);
    if (tid < 128) {
        TH[base_hi + 128 * 32 + xo[0]] = 1.0f;          // ones row behind S: db2 is column 64 of the dW2 accumulator
        Yhi[base_y + 7 * 32 + xo[7]] = 1.0f;            // ones row of x^T: db1 is column 7 of the tail accumulator
    }
    float* Phi = TH;                 // P: rows 0..63
    float* Plo = TL;
    float* Shi = TH + 64 * 32;       // S: rows 64..127
    float* Slo = TL + 64 * 32;
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(*tmem_slot);
    const uint32_t warp_u = uniform((uint32_t)warp);   // warp index the compiler knows to be uniform: single-UTCHMMA issue
    const int64_t n_tiles = (b.n + tcu::kRows - 1) / tcu::kRows;
    const int n_my = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // >= 1: the launcher sizes the grid to the tiles

    // ---- MMA-issuing warps ------------------------------------------------------------------------------------------------
    // One issuing thread sustains one tcgen05.mma per ~60 cycles whatever the shape (probes/tc_rate.cu); two issuing warps with
    // independent accumulator chains reach the tensor pipe's own rate (~33 cycles for M=128 N=64 K=8 tf32).
    if (warp_u >= 16) {
        const uint32_t sTH_ = smem_u32(TH), sTL_ = smem_u32(TL), sY_ = smem_u32(Yhi);
        const uint32_t sW2_ = smem_u32(smraw + tcu::oW2), sW2T_ = smem_u32(smraw + tcu::oW2T);
#pragma unroll 1
        for (int it = -1; it < n_my; ++it) {
            const bool has_cur = it >= 0, has_next = it + 1 < n_my;
            const uint32_t cur = (uint32_t)it & 1u, nxt = cur ^ 1u;
            const bool flush_now = it > 0 && (it % kFlushTiles) == 0;
            const uint32_t acc_w = (has_cur && (it % kFlushTiles) != 0) ? 1u : 0u;      // dW2|db2 accumulator restarts after a flush
            const uint32_t acc_t = (it == -1 || flush_now) ? 0u : 1u;                   // tail accumulator
            if (has_cur) {
                mbar_wait(&bars[tcu::RDY_0], cur);                   // completion index it
                fence_after_sync();
            }
            if (warp_u == 16) {
                if (has_cur && elect_one()) { issue_ts_64x64(tmem, tcu::cA + 128u * cur, sW2T_, tcu::cAcc + 64u * cur); mma_commit(&bars[tcu::BAR_D]); }   // dgrad(t)
                if (has_next) {
                    mbar_wait(&bars[tcu::RDY_1], nxt);               // completion index it+1
                    fence_after_sync();
                    if (elect_one()) { issue_ts_64x64(tmem, tcu::cA + 128u * nxt, sW2_, tcu::cAcc + 64u * nxt); mma_commit(&bars[tcu::BAR_FWD]); }  // fwd(u)
                }
                mbar_wait(&bars[tcu::RDY_2], nxt);                   // completion index it+1
                fence_after_sync();
                if (elect_one()) { issue_tail(tmem + tcu::cC2, sTH_, sTL_, sY_, acc_t, 2, 3); mma_commit(&bars[tcu::BAR_T]); }   // tail, hi*lo pass
            } else {
                if (has_cur && elect_one()) { issue_dw2(tmem + tcu::cW2, sTH_, sTL_, acc_w); mma_commit(&bars[tcu::BAR_W]); }    // dW2 | db2 (t)
                mbar_wait(&bars[tcu::RDY_2], nxt);                   // completion index it+1
                fence_after_sync();
                if (elect_one()) { issue_tail(tmem + tcu::cC, sTH_, sTL_, sY_, acc_t, 0, 2); mma_commit(&bars[tcu::BAR_T]); }    // dW1 | db1 | dWh
            }
        }
        return;
    }

    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);

    // normalisation constants and 1/B live in shared memory (read once per tile, uniformly): five fewer long-lived registers
    float* ncs = reinterpret_cast<float*>(red + PM_N) + 16;
    if (tid == 0) {
        float adv_mean = 0.f, adv_den = 1.f, ret_mean = 0.f, ret_den = 1.f;
        if (hp.normalize_adv) norm_consts(adv_mom, adv_mean, adv_den);
        if (hp.normalize_ret) norm_consts(ret_mom, ret_mean, ret_den);
        ncs[0] = adv_mean; ncs[1] = adv_den; ncs[2] = ret_mean; ncs[3] = ret_den; ncs[4] = 1.0f / (float)b.n;
    }
    tcu::compute_sync();

    // Metric partial sums and head-bias gradients: quantity q (q < PM_N: metric partial q, PM_N + r: bias gradient r) is
    // reduced by the chunk-(q & 3) warp of each row quadrant (all four compute the row's loss anyway): after a transposing
    // butterfly over the warp's 32 rows, lane q >> 2 keeps it -> ONE register instead of 26, the work spread over all 16 warps.
    float macc = 0.f;
    float zs0 = 0.f, zq0 = 0.f, zs1 = 0.f, zq1 = 0.f;
    uint32_t dead0 = 0, dead1 = 0;          // dead-unit counts of this warp's rows: lane i holds neuron 16*chunk + i

    float* out = grad_partials + (size_t)blockIdx.x * pstride;    // this CTA's partial gradient vector (16 B aligned)
    // Software-pipelined gather (stage s is issued by the row's chunk-s thread), register-free: cp.async copies the sample of the tile after next straight
    // into a per-row staging slot in shared memory ([x0..x7, action, logp_old, value_old, adv, ret, offset of the following
    // tile]); the owner reads its own slot one iteration later (cp.async.wait_all, no barrier: same thread).  Loads held in
    // registers across the iteration get spilled by ptxas, which turns every prefetch into a synchronous load.
    // The copies are issued in four small bursts spread over the iteration (stage 0..3): a single burst of all 7 scattered
    // loads of 128 rows (~900 sectors in flight) backs up the SM's L1TEX queue and stalls every warp's shared-memory loads
    // behind it for ~4000 cycles (measured with the clock64 trace); 128-256 sectors at a time do not.
    float* stg = reinterpret_cast<float*>(smraw + tcu::oStage) + row * 16;
    const uint32_t stg_s = smem_u32(stg);
    uint32_t pf_off = 0;                    // translated offset of the tile being prefetched
    float* sc1 = reinterpret_cast<float*>(smraw + tcu::oScal1) + row * 5 - 8;   // indexed like stg: scalars at [8..12]
    const uint32_t sc1_s = smem_u32(sc1);
    uint32_t* offn = reinterpret_cast<uint32_t*>(smraw + tcu::oOffN) + row;
    const bool packed = b.packed != nullptr;   // 64-byte sample records (gs_rollout_pack): the row is ONE aligned access
    // packed gather: each of the row's four chunk threads copies its 16-byte quarter of the record; chunk 0 also the next offset
    auto prefetch_record = [&](int64_t t_next) {
        const int64_t p = t_next * tcu::kRows + row;
        if (t_next < n_tiles && p < b.n) cp_async16(stg_s + 16 * chunk, b.packed + (int64_t)pf_off * GS_RECORD_FLOATS + 4 * chunk);
        else { stg[4 * chunk] = 0.f; stg[4 * chunk + 1] = 0.f; stg[4 * chunk + 2] = 0.f; stg[4 * chunk + 3] = 0.f; }
        if (chunk == 0) {
            const int64_t p2 = (t_next + gridDim.x) * tcu::kRows + row;
            if (t_next + gridDim.x < n_tiles && p2 < b.n) cp_async4(smem_u32(offn), offs + p2); else *offn = 0u;
        }
    };
    auto prefetch_stage = [&](int stage, int64_t t_next, uint32_t par = 0) {   // par: parity of the tile's position in this CTA's sequence
        float* sc = par ? sc1 : stg;
        const uint32_t sc_s = par ? sc1_s : stg_s;
        const int64_t p = t_next * tcu::kRows + row;
        const bool ok = t_next < n_tiles && p < b.n;
        const int64_t off = (int64_t)pf_off;
        if (stage == 0) {
            if (ok) {
                const float* o = b.obs + off * b.D;
                if (D4) {
                    if (b.D == 4) cp_async16(stg_s, o);
                    else { cp_async8(stg_s, o); stg[2] = 0.f; stg[3] = 0.f; }
                } else {
#pragma unroll
                    for (int d = 0; d < 8; ++d) { if (d < b.D) cp_async4(stg_s + 4 * d, o + d); else stg[d] = 0.f; }
                }
            } else {
#pragma unroll
                for (int d = 0; d < 8; ++d) stg[d] = 0.f;
            }
            const int64_t p2 = (t_next + gridDim.x) * tcu::kRows + row;
            if (t_next + gridDim.x < n_tiles && p2 < b.n) cp_async4(smem_u32(offn), offs + p2); else *offn = 0u;
        } else if (stage == 1) {
            if (ok) { cp_async4(sc_s + 32, b.actions + off); cp_async4(sc_s + 36, b.logp_old + off); }
            else { sc[8] = 0.f; sc[9] = 0.f; }
        } else if (stage == 2) {
            if (ok) { cp_async4(sc_s + 44, b.adv + off); cp_async4(sc_s + 48, b.ret + off); }
            else { sc[11] = 0.f; sc[12] = 0.f; }
        } else {
            if (ok && ALGO == ALGO_PPO) cp_async4(sc_s + 40, b.values_old + off); else sc[10] = 0.f;
        }
    };
    {
        const int64_t p0 = (int64_t)blockIdx.x * tcu::kRows + row;
        pf_off = p0 < b.n ? __ldg(offs + p0) : 0u;
        if (packed) prefetch_record(blockIdx.x);
        else {
#pragma unroll
            for (int st = 0; st < 4; ++st) if (chunk == st) prefetch_stage(st, blockIdx.x, 0u);
        }
    }

    float g[4] = {0.f, 0.f, 0.f, 0.f};      // d(loss)/d(head outputs) of the tile entering its backward phase
    float xk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // its observation (layer 1 is recomputed in the backward phase)
    uint32_t relu_mask2 = 0;                // relu: act'(h2) of the chunk, kept from the forward phase for the backward phase
#pragma unroll 1
    for (int it = -1; it < n_my; ++it) {
        const bool has_cur = it >= 0, has_next = it + 1 < n_my;
        const uint32_t cur = (uint32_t)it & 1u, nxt = cur ^ 1u;
        const uint32_t ph = cur;                                    // parity of the current tile's BAR_D / BAR_W completions
        const uint32_t a_cur = tcu::cA + 128u * cur, a_nxt = tcu::cA + 128u * nxt;
        const uint32_t acc_cur = tcu::cAcc + 64u * cur, acc_nxt = tcu::cAcc + 64u * nxt;
        const int64_t tile_u = blockIdx.x + (int64_t)(it + 1) * gridDim.x;
        const bool valid_u = has_next && (tile_u * tcu::kRows + row < b.n);
        const bool flush_now = it > 0 && (it % kFlushTiles) == 0;
        uint32_t relu_mask = 0;             // ACT == relu: act'(h1) of the chunk; tanh: d1[]
        float d1[ACT == GS_ACT_RELU ? 1 : 16];
        // ---- B1(t): dz2 -> TMEM A[cur] and P, h1^T -> S, x^T -> Y ---------------------------------------------------------------
        GS_TR(0);
        if (has_cur) {
            float dz[16];
            {
                float z[16];
                if (ACT != GS_ACT_RELU) {                            // tanh: act'(h2) needs h2 again (relu: the mask kept from F2)
                    tmem_ld16(lane_addr + acc_cur + 16 * chunk, z);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 w = *reinterpret_cast<const float4*>(whs + r * 64 + 16 * chunk + i);
                        d.x = fmaf(g[r], w.x, d.x); d.y = fmaf(g[r], w.y, d.y); d.z = fmaf(g[r], w.z, d.z); d.w = fmaf(g[r], w.w, d.w);
                    }
                    if (ACT == GS_ACT_RELU) {
                        dz[i] = ((relu_mask2 >> i) & 1u) ? d.x : 0.f;
                        dz[i + 1] = ((relu_mask2 >> (i + 1)) & 1u) ? d.y : 0.f;
                        dz[i + 2] = ((relu_mask2 >> (i + 2)) & 1u) ? d.z : 0.f;
                        dz[i + 3] = ((relu_mask2 >> (i + 3)) & 1u) ? d.w : 0.f;
                    } else {
                        const float4 bb = *reinterpret_cast<const float4*>(b2s + 16 * chunk + i);
                        dz[i] = d.x * act_bwd(act_fwd(z[i] + bb.x, ACT), ACT);
                        dz[i + 1] = d.y * act_bwd(act_fwd(z[i + 1] + bb.y, ACT), ACT);
                        dz[i + 2] = d.z * act_bwd(act_fwd(z[i + 2] + bb.z, ACT), ACT);
                        dz[i + 3] = d.w * act_bwd(act_fwd(z[i + 3] + bb.w, ACT), ACT);
                    }
                }
            }
            chunk_to_tmem(lane_addr + a_cur, chunk, dz);             // A[cur] is free: fwd(t) completed
            GS_TR(16);

            float z[16];
            layer1_chunk<D4>(w1s, b1s, xk, chunk, z);                // h1(t) again: cheaper than keeping it for a whole tile
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                z[i] = act_fwd(z[i], ACT);
                if (ACT == GS_ACT_RELU) relu_mask |= (z[i] > 0.f ? 1u : 0u) << i;
                else d1[ACT == GS_ACT_RELU ? 0 : i] = 1.0f - z[i] * z[i];
            }
            GS_TR(19);
            // the previous tail group reads P, S and Y: it must be done before they are rewritten
            mbar_wait(&bars[tcu::BAR_T], ph);                        // completion index it (iteration -1 issued #0)
            fence_after_sync();
            if (flush_now) flush_wgrad(lane_addr, quad, chunk, lane, row, m.D, A, m.has_value, out, it == kFlushTiles);
            GS_TR(17);
            chunk_to_transposed(Phi, Plo, base_hi, base_lo, xo, chunk, dz);
            GS_TR(18);
            chunk_to_transposed(Shi, Slo, base_hi, base_lo, xo, chunk, z);
            if (row_owner) {
#pragma unroll
                for (int d = 0; d < 7; ++d) {
                    const int idx = base_y + d * 32 + xo[d];
                    const float xh = tf32_rn(xk[d]);
                    Yhi[idx] = xh; Ylo[idx] = xk[d] - xh;
                }
            }
        }
        // ---- the gathered sample of tile u must have landed before the row's chunk threads read the staging slot -------------------
        if (has_next) cp_async_wait_all();                           // this thread's share of tile u's copies (issued an iteration ago) has landed
        GS_TR(1);
        tmem_st_wait();
        fence_proxy_async();
        fence_before_sync();
        tcu::compute_sync();                                                                                    // sync 1
        if (has_cur && tid == 0) tcu::mbar_arrive(&bars[tcu::RDY_0]);    // -> dgrad(t)
        fence_after_sync();
        GS_TR(2);
        // ---- F1(u): layer 1 -> h1 -> TMEM A[nxt] -------------------------------------------------------------------------------
        if (has_next) {
            float sx[16];                                            // the row's staging slot, read by its four chunk threads
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((D4 && q == 1) || q >= 2) continue;              // the sample scalars (8..12) are read in F3
                const float4 v4 = *reinterpret_cast<const float4*>(stg + 4 * q);
                sx[4 * q] = v4.x; sx[4 * q + 1] = v4.y; sx[4 * q + 2] = v4.z; sx[4 * q + 3] = v4.w;
            }
            if (D4) { sx[4] = sx[5] = sx[6] = sx[7] = 0.f; }
            pf_off = *offn;                                          // offset of the tile after u: its copies start after sync 2
            if (packed && chunk == 0) {                              // the record slot is rewritten after sync 2: F3 reads the scalars from sc1
#pragma unroll
                for (int d = 8; d < 13; ++d) sc1[d] = stg[d];
            }
            GS_TR(20);
#pragma unroll
            for (int d = 0; d < 8; ++d) xk[d] = sx[d];
            float z[16];
            GS_TR(21);
            layer1_chunk<D4>(w1s, b1s, xk, chunk, z);
            GS_TR(22);
            if (TRACK) chunk_stats(z, valid_u, lane, zs0, zq0, dead0);
            GS_TR(23);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = act_fwd(z[i], ACT);
            chunk_to_tmem(lane_addr + a_nxt, chunk, z);              // A[nxt] is free: dgrad(t-1) completed
        }
        GS_TR(4);
        tmem_st_wait();
        fence_before_sync();
        tcu::compute_sync();                                                                                    // sync 2
        if (has_next && tid == 0) tcu::mbar_arrive(&bars[tcu::RDY_1]);   // -> fwd(u), dW2(t)
        fence_after_sync();
        GS_TR(7);
        if (has_next) {                                              // every chunk thread has read the staging slots (sync 2)
            if (packed) prefetch_record(tile_u + gridDim.x);
            else if (chunk == 0) prefetch_stage(0, tile_u + gridDim.x);
        }
        // ---- dz1(t) = dh1 * act'(h1) -> dz1^T -> P, while fwd(u) runs ---------------------------------------------------------------
        if (has_cur) {
            float dz[16];
            mbar_wait(&bars[tcu::BAR_D], ph);
            fence_after_sync();
            tmem_ld16(lane_addr + acc_cur + 16 * chunk, dz);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (ACT == GS_ACT_RELU) dz[i] = ((relu_mask >> i) & 1u) ? dz[i] : 0.f;
                else dz[i] *= d1[ACT == GS_ACT_RELU ? 0 : i];
            }
            GS_TR(13);
            mbar_wait(&bars[tcu::BAR_W], ph);                        // dW2 has consumed dz2^T and h1^T: P and S may be overwritten
            fence_after_sync();
            GS_TR(14);
            chunk_to_transposed(Phi, Plo, base_hi, base_lo, xo, chunk, dz);
        }
        // ---- F2(u): z2 -> h2 -> this chunk's share of the head outputs -> TMEM scratch ------------------------------------------
        if (has_next) {
            mbar_wait(&bars[tcu::BAR_FWD], nxt);                     // completion index it+1
            fence_after_sync();
            GS_TR(9);
            float z[16];
            tmem_ld16(lane_addr + acc_nxt + 16 * chunk, z);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] += b2s[16 * chunk + i];
            if (TRACK) chunk_stats(z, valid_u, lane, zs1, zq1, dead1);
            if (ACT == GS_ACT_RELU) relu_mask2 = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                z[i] = act_fwd(z[i], ACT);
                if (ACT == GS_ACT_RELU) relu_mask2 |= (z[i] > 0.f ? 1u : 0u) << i;
            }
            chunk_to_transposed(Shi, Slo, base_hi, base_lo, xo, chunk, z);   // S is free: this thread waited for dW2(t) above
            float o4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(whs + r * 64 + 16 * chunk + i);
                    o4[r] = fmaf(z[i], w.x, o4[r]); o4[r] = fmaf(z[i + 1], w.y, o4[r]);
                    o4[r] = fmaf(z[i + 2], w.z, o4[r]); o4[r] = fmaf(z[i + 3], w.w, o4[r]);
                }
            }
            tmem_st4(lane_addr + tcu::cOP + 4 * chunk, o4);
            if (chunk == 1 && !packed) prefetch_stage(1, tile_u + gridDim.x, cur);
        }
        GS_TR(10);
        tmem_st_wait();
        fence_before_sync();
        tcu::compute_sync();                                                                                    // sync 3
        fence_after_sync();
        if (has_next && chunk == 2 && !packed) prefetch_stage(2, tile_u + gridDim.x, cur);
        GS_TR(11);
        // ---- F3(u): loss and d(loss)/d(heads), redundantly by the four chunk threads of the row (metrics: the owner) -------------
#pragma unroll
        for (int r = 0; r < 4; ++r) g[r] = 0.f;
        float pl_keep[PM_N];                                         // this row's metric terms (dead code outside the metric owners)
#pragma unroll
        for (int i = 0; i < PM_N; ++i) pl_keep[i] = 0.f;
        if (has_next) {
            float op[16];
            tmem_ld16(lane_addr + tcu::cOP, op);
            tmem_ld_wait();
            if (valid_u) {
                float outv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) outv[r] = bhs[r] + ((op[r] + op[4 + r]) + (op[8 + r] + op[12 + r]));
                float pl[PM_N];                                      // this sample's metric terms (non-owners: dead code)
#pragma unroll
                for (int i = 0; i < PM_N; ++i) pl[i] = 0.f;
                const float* sc = (packed || nxt) ? sc1 : stg;       // tile u's scalars: buffer of its parity (packed: always the copy)
                sample_loss<ALGO>(outv, A, __float_as_int(sc[8]), sc[9], sc[10], sc[11], sc[12], hp, ncs[0], ncs[1], ncs[2], ncs[3], ncs[4], g, pl);
#pragma unroll
                for (int i = 0; i < PM_N; ++i) pl_keep[i] = pl[i];
            }
        }
        if (has_next) {
            // Quantity q = 4 j + chunk (q < PM_N: metric term q of this row, PM_N + r: g[r]) is value j of this warp.  The 8 values
            // are reduced over the warp's 32 rows by a transposing butterfly: each xor level halves the number of live values
            // (a lane keeps the half its lane bit selects and sends the other), 4 + 2 + 1 + 1 + 1 = 9 shuffles in 5 dependent
            // levels instead of 7 x 5; lane L ends with the warp total of value L & 7.
            float v[8];
            auto pick = [&](auto ctag) {
                constexpr int Cc = decltype(ctag)::value;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = 4 * j + Cc;                            // a constant after unrolling
                    float x = 0.f;
                    if (q < PM_N) x = pl_keep[q < PM_N ? q : 0];
                    else if (q < PM_N + 4) x = g[(q - PM_N) & 3];
                    v[j] = x;
                }
            };
            if (chunk == 0) pick(std::integral_constant<int, 0>{});
            else if (chunk == 1) pick(std::integral_constant<int, 1>{});
            else if (chunk == 2) pick(std::integral_constant<int, 2>{});
            else pick(std::integral_constant<int, 3>{});
            const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
            float w4[4], u2[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float keep = b0 ? v[2 * j + 1] : v[2 * j], send = b0 ? v[2 * j] : v[2 * j + 1];
                w4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float keep = b1 ? w4[2 * j + 1] : w4[2 * j], send = b1 ? w4[2 * j] : w4[2 * j + 1];
                u2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            float t = (b2 ? u2[1] : u2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? u2[0] : u2[1], 4);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 16);
            macc += t;
        }
        if (row_owner) {                                             // Y is free: every thread waited for the previous tail in B1
#pragma unroll
            for (int r = 0; r < 4; ++r) {                            // zeros when there is no next tile: the tail's heads block adds 0
                const int idx = base_y + (8 + r) * 32 + xo[r];
                const float gh = tf32_rn(g[r]);
                Yhi[idx] = gh; Ylo[idx] = g[r] - gh;
            }
        }
        if (has_next && chunk == 3 && !packed) prefetch_stage(3, tile_u + gridDim.x, cur);
        GS_TR(12);
        GS_TR(15);
        fence_proxy_async();
        fence_before_sync();
        tcu::compute_sync();                                                                                    // sync 4
        if (tid == 0) tcu::mbar_arrive(&bars[tcu::RDY_2]);               // -> tail(t,u)
        fence_after_sync();
    }
    // ---- drain: wait for the last tail group and flush what the TMEM accumulators still hold --------------------------------
    mbar_wait(&bars[tcu::BAR_T], (uint32_t)n_my & 1u);               // completion index n_my
    fence_after_sync();
    flush_wgrad(lane_addr, quad, chunk, lane, row, m.D, A, m.has_value, out, n_my <= kFlushTiles);
    // ---- block reductions through shared-memory atomics: head biases (4 floats) and the metric partials (PM_N doubles) ------
    float* fr = reinterpret_cast<float*>(red + PM_N);            // [4 quadrants][4 heads] bias-gradient partials
    if (lane < 8 && 4 * lane + chunk < PM_N + 4) {               // lane L holds quantity 4 L + chunk (see F3)
        const int q = 4 * lane + chunk;
        if (q < PM_N) { if (macc != 0.f) atomicAdd(red + q, (double)macc); }
        else fr[quad * 4 + (q - PM_N)] = macc;
    }
    if (TRACK) {
        const double a0 = warp_sum((double)zs0), a1 = warp_sum((double)zq0), a2 = warp_sum((double)zs1), a3 = warp_sum((double)zq1);
        if (lane == 0) { atomicAdd(red + PM_Z0, a0); atomicAdd(red + PM_Z0SQ, a1); atomicAdd(red + PM_Z1, a2); atomicAdd(red + PM_Z1SQ, a3); }
        if (lane < 16) {
            if (dead0) atomicAdd(dead + 16 * chunk + lane, dead0);
            if (dead1) atomicAdd(dead + 64 + 16 * chunk + lane, dead1);
        }
    }
    fence_before_sync();
    tcu::compute_sync();
    if (tid < PM_N) metric_partials[(size_t)blockIdx.x * PM_N + tid] = red[tid];
    if (tid < 4) {
        const ParamOffsets po = param_offsets(m.D, 64, 64, m.A, m.has_value);
        const float sgm = (fr[tid] + fr[4 + tid]) + (fr[8 + tid] + fr[12 + tid]);     // fixed order: deterministic
        if (tid < A) out[po.bp + tid] = sgm;
        else if (tid == A && m.has_value) out[po.bv] = sgm;
    }
    if (warp == 0) tmem_dealloc(tmem, tcu::kTmemCols);
}

// ---- host launcher (called from update_kernels.cu::launch_update for the 64x64 network) ----------------------------------------
template <int ALGO>
int launch_update_tc(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                     const uint32_t* offs, float* grad_partials, int64_t pstride, double* metric_partials, uint32_t* dead, int grid, cudaStream_t st) {
    const bool d4 = md.D <= 4 && (md.D == 4 || md.D == 2);
    using KernelFn = void (*)(MlpDev, BatchDev, HpDev, const double*, const double*, const uint32_t*, float*, int64_t, double*, uint32_t*);
    auto pick_act = [&](auto act_tag) -> KernelFn {
        constexpr int ACT = decltype(act_tag)::value;
        if (track) return d4 ? update_tc_kernel<ALGO, true, true, ACT> : update_tc_kernel<ALGO, true, false, ACT>;
        return d4 ? update_tc_kernel<ALGO, false, true, ACT> : update_tc_kernel<ALGO, false, false, ACT>;
    };
    auto pick = [&]() -> KernelFn {
        if (md.act == GS_ACT_RELU) return pick_act(std::integral_constant<int, GS_ACT_RELU>{});
        return pick_act(std::integral_constant<int, GS_ACT_TANH>{});
    };
    auto kern = pick();
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tcu::kSmemBytes));
    kern<<<grid, tcu::kT, tcu::kSmemBytes, st>>>(md, b, hp, adv_mom, ret_mom, offs, grad_partials, pstride, metric_partials, dead);
    GS_LAUNCH_CHECK();
    return 0;
}
template int launch_update_tc<ALGO_PPO>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, const uint32_t*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);
template int launch_update_tc<ALGO_REINFORCE>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, const uint32_t*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);

}  // namespace gs

#ifdef GS_TC_TRACE
extern "C" int gs_debug_tc_trace(long long* host_out /* [18][24] */) {
    return (int)cudaMemcpyFromSymbol(host_out, gs::g_tc_trace, sizeof(long long) * 432);
}
#endif
