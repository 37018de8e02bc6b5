// update_tc.cu — tcgen05 / TMEM version of the fused minibatch update for the 64x64 MLP (PPO and REINFORCE).
//
// Same contract as update_kernels.cu::update_kernel (gather -> forward -> loss + metrics -> backward -> per-CTA partial
// gradients), but the three 64x64 GEMMs and every cross-sample reduction run on the 5th-gen tensor cores:
//
//   forward  z2[s][j]   = sum_k h1[s][k]  W2[j][k]        M=128 (samples)  N=64  K=64   A: TMEM (h1 rows),  B: W2   smem
//   dgrad    dh1[s][k]  = sum_j dz2[s][j] W2[j][k]        M=128            N=64  K=64   A: TMEM (dz2 rows), B: W2^T smem
//   wgrad    dW2[j][k] += sum_s dz2[s][j] h1[s][k]        M=64   N=64  K=128 (samples)  A: dz2^T smem, B: h1^T smem
//   heads    dWh[r][k] += sum_s g[s][r]   h2[s][k]        M=64 (k)  N=8 (r)  K=128      A: h2^T  smem, B: g^T  smem
//   layer 1  dW1[j][d] += sum_s dz1[s][j] x[s][d], db1    M=64   N=8 (d, col 7 = ones)  A: dz1^T smem, B: x^T  smem
//   bias 2   db2[j]    += sum_s dz2[s][j]                 M=64   N=8 (col 7 = ones)     A: dz2^T smem, B: x^T  smem
//
// Precision: every product is a 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM): ~1.5e-6 relative,
// measured on hardware by csrc/probes/tc_probe.cu, which keeps the 1e-4 gradient parity bar of the fp32 path.
// Layer 1, heads, softmax, loss and the activation derivatives are row-local fp32 SIMT work on registers; wgrad accumulators live in TMEM and are folded into fp32 partials every kFlushTiles tiles.
// All shared-memory operands use the K-major SWIZZLE_128B slab layout of tc_common.cuh (the only layout tf32 needs here);
// transposed copies ([feature][sample]) are written straight from registers, conflict-free (a warp writes 128 contiguous
// bytes of one row).  1 CTA of 512 threads per SM (thread = sample row x 16-column chunk), 225 KB shared memory, all 512 TMEM
// columns; MMA groups are issued by four different warps (one accumulator each, so every accumulation order is fixed).
#include <type_traits>

#include "mlp_tile.cuh"
#include "tc_common.cuh"
#include "update_shared.cuh"

namespace gs {

using namespace tc;

namespace tcu {
constexpr int kRows = 128;              // samples per tile == TMEM lanes
constexpr int kT = 512;                 // threads: thread = (row, 16-column chunk); warp w -> rows 32*(w&3).., chunk w>>2
// shared memory map (bytes)
constexpr int oQ = 0;                   // h1^T then dz1^T   hi [0,32K) lo [32K,64K)     [64 rows][128 samples]
constexpr int oR = 65536;               // h2^T then dz2^T
constexpr int oW2 = 131072;             // W2  hi 16K, lo 16K      [64 j][64 k]
constexpr int oW2T = 163840;            // W2^T                    [64 k][64 j]
constexpr int oGT = 196608;             // g^T  hi 4K, lo 4K       [8 r][128 samples]
constexpr int oXT = 204800;             // x^T  hi 4K, lo 4K       [8 d][128 samples], row 7 = ones
constexpr int oMisc = 212992;
constexpr int oW1 = oMisc;              // [64][8] fp32
constexpr int oB1 = oW1 + 2048;
constexpr int oB2 = oB1 + 256;
constexpr int oWH = oB2 + 256;          // [4][64]
constexpr int oBH = oWH + 1024;         // [4]
constexpr int oBar = oBH + 16;          // 5 mbarriers
constexpr int oTmem = oBar + 64;
constexpr int oRed = oTmem + 16;        // PM_N doubles + 4 floats (block reductions through shared atomics)
constexpr int oOutP = oRed + 256;       // partial head outputs [4 chunks][128 rows] float4
constexpr int kSmemBytes = oOutP + 4 * 128 * 16;
static_assert(kSmemBytes <= 232448, "shared memory budget");
// TMEM columns
constexpr uint32_t cAhi = 0, cAlo = 64, cAcc = 128, cW2 = 192, cWH = 256, cW1 = 264, cB2 = 272, kTmemCols = 512;
enum { BAR_FWD = 0, BAR_H, BAR_D, BAR_W, BAR_1 };
}  // namespace tcu

// operands are stored as (hi, lo) = (rn_tf32(x), x - hi): see tc_common.cuh::tf32_rn

// 16 values of the thread's row chunk -> TMEM A columns (hi = full fp32: the tensor core ignores the low 13 bits; lo = rest)
__device__ __forceinline__ void chunk_to_tmem(uint32_t lane_addr, int c, const float (&v)[16]) {
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { h[i] = tf32_rn(v[i]); l[i] = v[i] - h[i]; }
    tmem_st16(lane_addr + tcu::cAhi + 16 * c, h);
    tmem_st16(lane_addr + tcu::cAlo + 16 * c, l);
}
// the chunk -> column s of a [64 rows][128 samples] K-major swizzled tile (rows 16c .. 16c+15), hi and lo copies.
// A warp writes 128 contiguous bytes per row: conflict-free.
__device__ __forceinline__ void chunk_to_transposed(float* hi, float* lo, int base_s, const int (&xo)[8], int c, const float (&v)[16]) {
    const int b = base_s + c * 512;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = b + i * 32 + xo[i & 7];
        const float h = tf32_rn(v[i]);
        hi[idx] = h;
        lo[idx] = v[i] - h;
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
// The dead test is one FMNMX per element; the (rare) counting path runs only when the chunk minimum trips it.
__device__ __forceinline__ void chunk_stats(const float (&z)[16], float& s, float& q, uint32_t* dead_base) {
    float mn = 1.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s += z[i]; q = fmaf(z[i], z[i], q); mn = fminf(mn, fabsf(z[i])); }
    if (mn < 1e-6f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (fabsf(z[i]) < 1e-6f) atomicAdd(dead_base + i, 1u);
    }
}

// K-major SWIZZLE_128B descriptor with SBO = 1024 B, LBO = 16 B: only the 14-bit start-address field changes between MMAs
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    return ((uint64_t)(0x40004040u) << 32) | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | 0x10000u);
}

// 3xTF32 groups, each issued by ONE thread ------------------------------------------------------------------------------------
// A from TMEM (M=128 rows = lanes), B K-major smem [64 rows][64 k] (lo copy 16 KB after hi)
__device__ __forceinline__ void issue_ts_64x64(uint32_t tmem_base, uint32_t b_hi_addr, uint32_t acc_col) {
    const uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
    uint32_t accumulate = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t acol = tmem_base + (pass == 1 ? tcu::cAlo : tcu::cAhi);
        const uint64_t bd = desc_sw128(b_hi_addr + (pass == 2 ? 16384u : 0u));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            mma_tf32_ts(tmem_base + acc_col, acol + kk * 8, bd + (uint64_t)(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4), idesc, accumulate);
            accumulate = 1;
        }
    }
}
// A = [64 rows][128 samples] K-major smem (lo copy 32 KB after hi); B = [NB rows][128 samples] (lo copy at +b_lo_off)
template <int NB>
__device__ __forceinline__ void issue_ss_wgrad(uint32_t tmem_acc, uint32_t a_addr, uint32_t b_addr, uint32_t b_lo_off, uint32_t accumulate) {
    const uint32_t idesc = make_idesc_tf32(64, NB, 0, 0);
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint64_t ad = desc_sw128(a_addr + (pass == 1 ? 32768u : 0u));
        const uint64_t bd = desc_sw128(b_addr + (pass == 2 ? b_lo_off : 0u));
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            mma_tf32(tmem_acc, ad + (uint64_t)(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4),
                     bd + (uint64_t)(((kk >> 2) * (NB * 128) + (kk & 3) * 32) >> 4), idesc, accumulate);
            accumulate = 1;
        }
    }
}

// Move the wgrad accumulators from TMEM into this CTA's fp32 partial-gradient vector (global, L2-resident) and let the next
// MMA group restart from zero.  The tensor core accumulates with truncation, so a TMEM chain is kept to kFlushTiles tiles
// (bias ~1e-5 relative, measured); across chains the sums are round-to-nearest fp32 adds in a fixed order (deterministic).
// M=64 accumulators: row m lives in lane (m/16)*32 + m%16, i.e. lanes 0..15 of the warps of quadrant m/16.
constexpr int kFlushTiles = 8;
__device__ __forceinline__ void flush_wgrad(uint32_t lane_addr, int quad, int chunk, int lane, int D, int A, int has_value,
                                            const ParamOffsets& po, float* __restrict__ out, bool first) {
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
    if (chunk < 2) tmem_ld16(lane_addr + (chunk == 0 ? tcu::cWH : tcu::cB2), s16);   // cols 256..271: dWh^T | dW1 ; 272..287: db2 in col 7
    tmem_ld_wait();
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
        auto put = [&](int64_t idx, float val) { out[idx] = first ? val : out[idx] + val; };
        if (chunk == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) if (r < A) put(po.wp + r * 64 + mrow, s16[r]);
            if (has_value) put(po.wv + mrow, A == 2 ? s16[2] : s16[3]);
#pragma unroll
            for (int d = 0; d < 7; ++d) if (d < D) put(po.w1 + mrow * D + d, s16[8 + d]);
            put(po.b1 + mrow, s16[15]);
        } else if (chunk == 1) {
            put(po.b2 + mrow, s16[7]);
        }
    }
}

template <int ALGO, bool TRACK, bool D4, int ACT>
__global__ void __launch_bounds__(tcu::kT, 1)
update_tc_kernel(MlpDev m, BatchDev b, HpDev hp, const double* __restrict__ adv_mom, const double* __restrict__ ret_mom,
                 float* __restrict__ grad_partials, int64_t pstride, double* __restrict__ metric_partials, uint32_t* __restrict__ dead) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* Qhi = reinterpret_cast<float*>(smraw + tcu::oQ);
    float* Qlo = Qhi + 8192;
    float* Rhi = reinterpret_cast<float*>(smraw + tcu::oR);
    float* Rlo = Rhi + 8192;
    float* GThi = reinterpret_cast<float*>(smraw + tcu::oGT);
    float* GTlo = GThi + 1024;
    float* XThi = reinterpret_cast<float*>(smraw + tcu::oXT);
    float* XTlo = XThi + 1024;
    float* w1s = reinterpret_cast<float*>(smraw + tcu::oW1);
    float* b1s = reinterpret_cast<float*>(smraw + tcu::oB1);
    float* b2s = reinterpret_cast<float*>(smraw + tcu::oB2);
    float* whs = reinterpret_cast<float*>(smraw + tcu::oWH);
    float* bhs = reinterpret_cast<float*>(smraw + tcu::oBH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + tcu::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smraw + tcu::oTmem);
    double* red = reinterpret_cast<double*>(smraw + tcu::oRed);
    float4* outp = reinterpret_cast<float4*>(smraw + tcu::oOutP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127;              // sample row of the tile == TMEM lane
    const int chunk = tid >> 7;             // 16-column chunk of the row this thread owns
    const int quad = warp & 3;              // TMEM lane quadrant of the warp
    const bool row_owner = chunk == 0;      // does the per-row scalar work (gather, loss)
    const int A = m.A;
    const ParamOffsets po = param_offsets(m.D, 64, 64, m.A, m.has_value);

    // ---- one-time staging ------------------------------------------------------------------------------------------------
    if (warp == 0) tmem_alloc(tmem_slot, tcu::kTmemCols);
    if (tid == 0) {
        mbar_init(&bars[tcu::BAR_FWD], 1); mbar_init(&bars[tcu::BAR_H], 1); mbar_init(&bars[tcu::BAR_D], 1);
        mbar_init(&bars[tcu::BAR_W], 2);   // dW2 and db2 groups are issued by two different warps
        mbar_init(&bars[tcu::BAR_1], 1);
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
        for (int i = tid; i < 1024; i += tcu::kT) { GThi[i] = 0.f; GTlo[i] = 0.f; XThi[i] = 0.f; XTlo[i] = 0.f; }
        if (tid < PM_N + 2) red[tid] = 0.0;
    }
    __syncthreads();
    // transposed-store constants of this thread (sample column s = row)
    const int chunk_s = (row & 31) >> 2;
    int xo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) xo[c] = (chunk_s ^ c) << 2;
    const int base64 = (row >> 5) * 2048 + (row & 3);   // [64 rows][128] tiles
    const int base8 = (row >> 5) * 256 + (row & 3);     // [8 rows][128] tiles
    if (row_owner) XThi[base8 + 7 * 32 + xo[7]] = 1.0f; // ones row: bias gradients fall out of the same MMAs
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t sQ = smem_u32(Qhi), sR = smem_u32(Rhi), sW2 = smem_u32(smraw + tcu::oW2), sW2T = smem_u32(smraw + tcu::oW2T);
    const uint32_t sGT = smem_u32(GThi), sXT = smem_u32(XThi);

    float adv_mean = 0.f, adv_den = 1.f, ret_mean = 0.f, ret_den = 1.f;
    if (hp.normalize_adv) norm_consts(adv_mom, adv_mean, adv_den);
    if (hp.normalize_ret) norm_consts(ret_mom, ret_mean, ret_den);
    const float invB = 1.0f / (float)b.n;

    float pm[PM_N];
#pragma unroll
    for (int i = 0; i < PM_N; ++i) pm[i] = 0.f;
    float gsum[4] = {0.f, 0.f, 0.f, 0.f};   // head bias gradients (sum over this thread's samples; row owners only)
    float zs0 = 0.f, zq0 = 0.f, zs1 = 0.f, zq1 = 0.f;

    float* out = grad_partials + (size_t)blockIdx.x * pstride;    // this CTA's partial gradient vector (16 B aligned)
    const int64_t n_tiles = (b.n + tcu::kRows - 1) / tcu::kRows;
    uint32_t it = 0;
    // software-pipelined gather: row owners load the next tile's sample while the current tile is being processed
    float nx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int n_a = 0;
    float n_lp = 0.f, n_v = 0.f, n_adv = 0.f, n_ret = 0.f;
    auto prefetch = [&](int64_t t_next) {
#pragma unroll
        for (int d = 0; d < 8; ++d) nx[d] = 0.f;
        n_a = 0; n_lp = n_v = n_adv = n_ret = 0.f;
        const int64_t p = t_next * tcu::kRows + row;
        if (row_owner && t_next < n_tiles && p < b.n) {
            const int64_t off = sample_offset(b, p);
            const float* o = b.obs + off * b.D;
            if (D4) {
                if (b.D == 4) { const float4 v4 = __ldg(reinterpret_cast<const float4*>(o)); nx[0] = v4.x; nx[1] = v4.y; nx[2] = v4.z; nx[3] = v4.w; }
                else { const float2 v2 = __ldg(reinterpret_cast<const float2*>(o)); nx[0] = v2.x; nx[1] = v2.y; }
            } else {
#pragma unroll
                for (int d = 0; d < 7; ++d) if (d < b.D) nx[d] = __ldg(o + d);
            }
            n_a = __ldg(b.actions + off);
            n_lp = __ldg(b.logp_old + off);
            n_adv = __ldg(b.adv + off);
            n_ret = __ldg(b.ret + off);
            if (ALGO == ALGO_PPO) n_v = __ldg(b.values_old + off);
        }
    };
    prefetch(blockIdx.x);
#pragma unroll 1
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int64_t pos = tile * tcu::kRows + row;
        const bool valid = pos < b.n;
        // ---- this tile's sample data was prefetched during the previous tile (row owners) --------------------------------------
        float x[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) x[d] = nx[d];
        const int a_s = n_a;
        const float lp_old = n_lp, v_old = n_v, adv_s = n_adv, ret_s = n_ret;
        prefetch(tile + gridDim.x);                 // global-load latency of the NEXT tile hides behind this tile's work
        // previous tile's last MMAs (dW1: reads Q and x^T) must be done before Q / x^T / TMEM-A are rewritten
        if (it > 0) {
            mbar_wait(&bars[tcu::BAR_1], ph ^ 1);
            fence_after_sync();
            if (it % kFlushTiles == 0) flush_wgrad(lane_addr, quad, chunk, lane, m.D, A, m.has_value, po, out, it == kFlushTiles);
        }
        const uint32_t acc_flag = (it % kFlushTiles) != 0 ? 1u : 0u;
        if (row_owner) {
#pragma unroll
            for (int d = 0; d < 7; ++d) {
                const int idx = base8 + d * 32 + xo[d];
                const float xh = tf32_rn(x[d]);
                XThi[idx] = xh; XTlo[idx] = x[d] - xh;
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        // ---- layer 1 chunk: h1 -> TMEM A (forward operand) and h1^T -> Q (wgrad operand) ----------------------------------------
        if (!row_owner) {
#pragma unroll
            for (int d = 0; d < 7; ++d) x[d] = XThi[base8 + d * 32 + xo[d]] + XTlo[base8 + d * 32 + xo[d]];
        }
        {
            float z[16];
            layer1_chunk<D4>(w1s, b1s, x, chunk, z);
            if (TRACK && valid) chunk_stats(z, zs0, zq0, dead + 16 * chunk);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = act_fwd(z[i], ACT);
            chunk_to_tmem(lane_addr, chunk, z);
            chunk_to_transposed(Qhi, Qlo, base64, xo, chunk, z);
        }
        tmem_st_wait();
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 0) { issue_ts_64x64(tmem, sW2, tcu::cAcc); mma_commit(&bars[tcu::BAR_FWD]); }
        // ---- layer 2 epilogue: z2 chunk from TMEM, h2^T -> R (head-wgrad operand), partial head dot products ---------------------
        mbar_wait(&bars[tcu::BAR_FWD], ph);
        fence_after_sync();
        {
            float z[16];
            tmem_ld16(lane_addr + tcu::cAcc + 16 * chunk, z);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] += b2s[16 * chunk + i];
            if (TRACK && valid) chunk_stats(z, zs1, zq1, dead + 64 + 16 * chunk);
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = act_fwd(z[i], ACT);
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
            outp[chunk * 128 + row] = make_float4(o4[0], o4[1], o4[2], o4[3]);
            chunk_to_transposed(Rhi, Rlo, base64, xo, chunk, z);   // R is free: BAR_1 of the previous tile implies its BAR_W
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        // ---- loss (row owners) ---------------------------------------------------------------------------------------------
        if (row_owner) {
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                const float4 p0 = outp[row], p1 = outp[128 + row], p2 = outp[256 + row], p3 = outp[384 + row];
                float outv[4];
                outv[0] = bhs[0] + ((p0.x + p1.x) + (p2.x + p3.x)); outv[1] = bhs[1] + ((p0.y + p1.y) + (p2.y + p3.y));
                outv[2] = bhs[2] + ((p0.z + p1.z) + (p2.z + p3.z)); outv[3] = bhs[3] + ((p0.w + p1.w) + (p2.w + p3.w));
                sample_loss<ALGO>(outv, A, a_s, lp_old, v_old, adv_s, ret_s, hp, adv_mean, adv_den, ret_mean, ret_den, invB, g, pm);
#pragma unroll
                for (int r = 0; r < 4; ++r) gsum[r] += g[r];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int idx = base8 + r * 32 + xo[r];
                const float gh = tf32_rn(g[r]);
                GThi[idx] = gh; GTlo[idx] = g[r] - gh;
            }
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 32) { issue_ss_wgrad<8>(tmem + tcu::cWH, sR, sGT, 4096u, acc_flag); mma_commit(&bars[tcu::BAR_H]); }
        // ---- dz2 chunk = (g . Wh) * act'(h2): h2 re-derived from the z2 accumulator still in TMEM --------------------------------
        float dz[16];
        {
            float g[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) g[r] = GThi[base8 + r * 32 + xo[r]] + GTlo[base8 + r * 32 + xo[r]];
            tmem_ld16(lane_addr + tcu::cAcc + 16 * chunk, dz);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 w = *reinterpret_cast<const float4*>(whs + r * 64 + 16 * chunk + i);
                    d.x = fmaf(g[r], w.x, d.x); d.y = fmaf(g[r], w.y, d.y); d.z = fmaf(g[r], w.z, d.z); d.w = fmaf(g[r], w.w, d.w);
                }
                const float4 bb = *reinterpret_cast<const float4*>(b2s + 16 * chunk + i);
                dz[i] = d.x * act_bwd(act_fwd(dz[i] + bb.x, ACT), ACT);
                dz[i + 1] = d.y * act_bwd(act_fwd(dz[i + 1] + bb.y, ACT), ACT);
                dz[i + 2] = d.z * act_bwd(act_fwd(dz[i + 2] + bb.z, ACT), ACT);
                dz[i + 3] = d.w * act_bwd(act_fwd(dz[i + 3] + bb.w, ACT), ACT);
            }
        }
        chunk_to_tmem(lane_addr, chunk, dz);            // TMEM A is free (forward MMAs completed)
        mbar_wait(&bars[tcu::BAR_H], ph);               // head wgrad has consumed h2^T: R may be overwritten
        fence_after_sync();
        chunk_to_transposed(Rhi, Rlo, base64, xo, chunk, dz);
        tmem_st_wait();
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 0) { issue_ts_64x64(tmem, sW2T, tcu::cAcc); mma_commit(&bars[tcu::BAR_D]); }                               // dgrad
        if (tid == 32) { issue_ss_wgrad<64>(tmem + tcu::cW2, sR, sQ, 32768u, acc_flag); mma_commit(&bars[tcu::BAR_W]); }      // dW2
        if (tid == 64) { issue_ss_wgrad<8>(tmem + tcu::cB2, sR, sXT, 4096u, acc_flag); mma_commit(&bars[tcu::BAR_W]); }       // db2
        // ---- dz1 chunk = dh1 * act'(h1) -> dz1^T -> Q -------------------------------------------------------------------------------
        {
            float h1[16];
            layer1_chunk<D4>(w1s, b1s, x, chunk, h1);    // overlaps the dgrad MMAs
            mbar_wait(&bars[tcu::BAR_D], ph);
            fence_after_sync();
            tmem_ld16(lane_addr + tcu::cAcc + 16 * chunk, dz);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) dz[i] *= act_bwd(act_fwd(h1[i], ACT), ACT);
        }
        mbar_wait(&bars[tcu::BAR_W], ph);               // wgrad has consumed h1^T: Q may be overwritten
        fence_after_sync();
        chunk_to_transposed(Qhi, Qlo, base64, xo, chunk, dz);
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 96) { issue_ss_wgrad<8>(tmem + tcu::cW1, sQ, sXT, 4096u, acc_flag); mma_commit(&bars[tcu::BAR_1]); }
    }
    // ---- drain: wait for the last MMAs and flush what the TMEM accumulators still hold -------------------------------------
    if (it > 0) {
        mbar_wait(&bars[tcu::BAR_1], (it - 1) & 1);
        fence_after_sync();
        flush_wgrad(lane_addr, quad, chunk, lane, m.D, A, m.has_value, po, out, it <= kFlushTiles);
    }
    // ---- block reductions through shared-memory atomics: head biases (4 floats) and the metric partials (PM_N doubles) ------
    float* fr = reinterpret_cast<float*>(red + PM_N);
    if (TRACK) { pm[PM_Z0] = zs0; pm[PM_Z0SQ] = zq0; pm[PM_Z1] = zs1; pm[PM_Z1SQ] = zq1; }
    if (row_owner || TRACK) {
#pragma unroll
        for (int i = 0; i < PM_N; ++i) {
            const double v = warp_sum((double)pm[i]);
            if (lane == 0 && v != 0.0) atomicAdd(red + i, v);
        }
    }
    if (row_owner) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float sgm = warp_sum(gsum[r]);
            if (lane == 0) atomicAdd(fr + r, sgm);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (tid < PM_N) metric_partials[(size_t)blockIdx.x * PM_N + tid] = red[tid];
    if (tid < 4) {
        if (tid < A) out[po.bp + tid] = fr[tid];
        else if (tid == A && m.has_value) out[po.bv] = fr[tid];
    }
    if (warp == 0) tmem_dealloc(tmem, tcu::kTmemCols);
}

// ---- host launcher (called from update_kernels.cu::launch_update for the 64x64 network) ----------------------------------------
template <int ALGO>
int launch_update_tc(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                     float* grad_partials, int64_t pstride, double* metric_partials, uint32_t* dead, int grid, cudaStream_t st) {
    const bool d4 = md.D <= 4 && (md.D == 4 || md.D == 2);
    using KernelFn = void (*)(MlpDev, BatchDev, HpDev, const double*, const double*, float*, int64_t, double*, uint32_t*);
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
    kern<<<grid, tcu::kT, tcu::kSmemBytes, st>>>(md, b, hp, adv_mom, ret_mom, grad_partials, pstride, metric_partials, dead);
    GS_LAUNCH_CHECK();
    return 0;
}
template int launch_update_tc<ALGO_PPO>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);
template int launch_update_tc<ALGO_REINFORCE>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);

}  // namespace gs
