// update_f16.cu — tcgen05 / TMEM fused minibatch update (PPO and REINFORCE) for the H x H MLPs, H = 64 and 128.
//
// Same contract as update_kernels.cu::update_kernel (gather -> forward -> loss + metrics -> backward -> per-CTA partial
// gradients; reference: agents/ppo/ppo_agent.py:21-152, agents/reinforce/reinforce_agent.py:11-88).  EVERY contraction of the
// step runs on the 5th-gen tensor cores as fp16x3 (`kind::f16`, K = 16): a value x is carried as hi = fp16(x), lo = fp16(x - hi)
// (22 significant bits), a product as lo*hi + hi*lo + hi*hi with fp32 accumulation in TMEM.  fp16 rather than bf16 pairs (16 bits):
// kind::f16 takes ONE format for both operands (probes/mma16_probe.cu: mixed f16 x bf16 descriptors fault), and with bf16 pairs the
// per-sample head outputs carry ~1e-5 of error, which flips samples across the kinks of the PPO loss (clip range, clipped value loss):
// measured 1.1e-3 relative gradient error on a 1 M-sample minibatch against 1.5e-6 for fp32 (profiles/r2_operand_format.md).  fp16's
// range is kept safe by carrying the BACKWARD quantities (g, dz2, dz1) for the SUM loss (x B, O(advantage) instead of O(1/B)); the 1/B
// goes into the fold of the accumulators.  Values below 6e-5 lose relative, not absolute, precision (error <= 3e-8); values above
// 65504 (never reached by the classic-control observations, their activations or returns) saturate.
//
// ONE physical layout for every activation tile: [128 samples][H] fp16 in 64-column slabs of 128-byte rows with the
// SWIZZLE_128B chunk permutation (16-byte chunk c of row r at chunk c ^ (r & 7)).  The same bytes are
//   * a K-major  operand with MN = sample, K = feature   (forward, dgrad:  A operand), and
//   * an MN-major operand with MN = feature, K = sample  (weight gradients: A and B operand),
// so an activation is written once, row-major, with 16-byte stores by the thread that owns the row (probes/mma16_probe.cu
// holds the known-answer tests of every descriptor form used here).  Weights are staged once per launch the same way:
// W2[j][k] is K-major B for the forward pass and MN-major B for dgrad.
//
// Small operands ride in a per-set "X" tile [128 samples][64] of four 16-column groups:
//   group 0  g16 = [g_hi(4) g_lo(4) g_hi(4) 0 0 1 1]   d(loss)/d(head outputs) of the tile + a pair of ones
//   group 1/2  x16 = [x_hi(D) . x_lo(D) . 1 1] (cols 0.., 7.., 14, 15) of even / odd tiles (cp.async straight from the packed record)
//   group 3  the record's scalars {action, logp_old, value_old, adv, ret} (not an operand)
// and in "WS" [H][64]: group 0 = [W1_hi W1_hi b1_hi b1_lo], 1 = [W1_lo], 2 = [b2_hi b2_lo at 14, 15], 3 = head rows
// [Wh_hi(4) Wh_hi(4) Wh_lo(4) 0(4)]^T.  With them layer 1, every bias, the head outputs and d(loss)/dh2 are single K = 16
// MMAs, and every bias gradient is a column of a weight-gradient accumulator:
//   L1     z1  = x16 . WS0 + x16 . WS1                                   (2 MMAs)          -> acc0
//   fwd    z2  = x16 . WS2 (bias) + h1 . W2^T   (fp16x3)                 (1 + 3 H/16)      -> acc1
//   heads  out = h2 . WS3 (MN-major B, N = 16; 2 passes; out[r] = col r + col 8+r)        -> accH
//   dh2        = g16 . WS3 (K-major B)                                   (1)               -> acc0
//   W-c    dWh^T[k][c] += sum_s h2[s][k] g16[s][c]      (MN-major A = h2, B = X group 0; dWh[r][k] = col r + col 4+r)
//   dgrad  dh1 = dz2 . W2 (MN-major B)                                                     -> acc1
//   W-a    [dW2 | db2][j][n] += sum_s dz2[s][j] [h1 | g16][s][n]   (N = H + 16: the X tile sits one slab stride behind
//          h1_hi, so LBO reaches it; db2 = column H + 14, the ones of g16)
//   W-b    [dW1 | db1][j][c] += sum_s dz1[s][j] x16[s][c]           (dW1[j][d] = col d + col 7+d, db1 = col 14)
//
// Schedule: one CTA per SM, 16 compute warps + one MMA-issuing warp per set.  A SET owns a 128-sample tile at a time and
// two tile buffers P (h1) and Q (h2, then dz2, then dz1); thread = (sample row, 32-column group).  H = 64 runs TWO sets
// on alternating tiles so that one set's SIMT stages overlap the other's MMA groups; H = 128 runs one set.  Per tile a set's
// compute warps walk five stages, each gated by an MMA-completion mbarrier (tcgen05.commit) and closed by a per-warp arrive
// on an operands-ready mbarrier (no CTA-wide barrier in the loop):
//   A  z1 -> relu -> h1 (hi, lo) -> P            B  z2 -> relu -> h2 -> Q
//   C  (one thread per row) out -> loss, metrics, g -> g16; cp.async of the next tile's record into X
//   D  dh2 * act'(h2) -> dz2 -> Q  (after W-c released Q)      E  dh1 * act'(h1 in P) -> dz1 -> Q  (after W-a released Q)
// act' is re-derived from the activation still held in its tile (relu: hi != 0), so nothing but two offsets lives across stages.
// The next tile's layer 1 is issued ahead of W-b, and W-b reads Q while stage A of the next tile already writes P.
// Weight-gradient accumulators stay in TMEM for the whole launch (folded into the set's fp32 partial vector every 64 tiles: the
// tensor core's accumulation truncates, but 28 tiles against a fold after every tile differ by 5e-6 of the gradient norm, measured;
// a mid-kernel fold costs ~7,000 cycles because the partial vector has left L2 by then); partial vectors are summed in a fixed
// order by gs_update_finish -> deterministic gradients.
#include <type_traits>

#define GS_FAST_TRANSCENDENTALS   // the loss stage is on the tile's critical path: ex2.approx / lg2.approx (2^-21 relative) instead of libm
#include "mlp_tile.cuh"
#include "f16x3.cuh"
#include "update_shared.cuh"

namespace gs {

using namespace tc;

namespace hfu {

#ifndef GS_BF_FLUSH
#define GS_BF_FLUSH 64
#endif
constexpr int kFlushTiles = GS_BF_FLUSH;

enum { BAR_Z1 = 0, BAR_Z2, BAR_OUT, BAR_DH2, BAR_WC, BAR_DH1, BAR_WA, BAR_WB, RDY_X, RDY_H1, RDY_H2, RDY_G, RDY_DZ2, RDY_DZ1, kBars };

template <int H>
struct Cfg {
    static constexpr int kSets = H == 64 ? 2 : 1;
    static constexpr int kCG = H / 32;                     // 32-column groups of a row
    static constexpr int kCW = 4 * kCG;                    // compute warps of a set
    static constexpr int kComputeWarps = kSets * kCW;      // 16
    static constexpr int kCompute = kComputeWarps * 32;
    static constexpr int kThreads = kCompute + 32 * kSets;
    static constexpr int kSlabs = H / 64;
    static constexpr uint32_t kTile = kSlabs * kSlab;      // one precision of one activation tile
    // per-set regions (X directly behind P_hi: it is "slab kSlabs" of the h1 operand)
    static constexpr uint32_t oPhi = 0, oX = kTile, oPlo = oX + kSlab, oQhi = oPlo + kTile, oQlo = oQhi + kTile, kSetBytes = oQlo + kTile;
    // shared by the sets
    static constexpr uint32_t kW2 = H * H * 2, kW2Slab = H * 128;
    static constexpr uint32_t oW2hi = kSets * kSetBytes, oW2lo = oW2hi + kW2, oWS = oW2lo + kW2, oMisc = oWS + H * 128;
    static constexpr uint32_t oBars = oMisc;                          // kSets * kBars mbarriers
    static constexpr uint32_t oTmem = oBars + 8 * kBars * kSets;
    static constexpr uint32_t oBH = oTmem + 16;                       // 4 head biases
    static constexpr uint32_t oNcs = oBH + 16;                        // 8 floats: normalisation constants, 1/B
    static constexpr uint32_t oRed = oNcs + 32;                       // PM_N doubles
    static constexpr uint32_t oFr = oRed + 8 * PM_N;                  // [kSets][4 quads][4] head-bias gradient partials
    static constexpr uint32_t kSmemBytes = oFr + 4 * 16 * kSets;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
    static_assert(oW2hi % 1024 == 0 && oWS % 1024 == 0 && kSetBytes % 1024 == 0, "swizzle atoms are 1024-byte aligned");
    // TMEM columns of a set
    static constexpr uint32_t cAcc0 = 0, cAcc1 = H, cW2 = 2 * H, cH = cW2 + H + 16, cW1 = cH + 16, cWh = cW1 + 16, kSetCols = cWh + 16;
    static_assert(kSets * kSetCols <= 512, "TMEM budget");
};

// the 64-byte sample record of the gather: {x16 (32 B), action, logp_old, value_old, adv, ret, 0, 0, 0}
__device__ __forceinline__ void make_record(const BatchDev& b, int64_t off, uint4 (&rec)[4]) {
    uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    auto put = [&](int e, uint32_t bits) { w[e >> 1] |= bits << ((e & 1) * 16); };
#pragma unroll
    for (int d = 0; d < kMaxD; ++d)
        if (d < b.D) {
            const float x = __ldg(b.obs + off * b.D + d);
            const float xh = f16_hi(x);
            put(d, f16_bits(x));
            put(7 + d, f16_bits(x - xh));
        }
    w[7] = kOnes2;   // ones at 14, 15
    rec[0] = make_uint4(w[0], w[1], w[2], w[3]);
    rec[1] = make_uint4(w[4], w[5], w[6], w[7]);
    rec[2] = make_uint4((uint32_t)__ldg(b.actions + off), __float_as_uint(__ldg(b.logp_old + off)),
                        b.values_old ? __float_as_uint(__ldg(b.values_old + off)) : 0u, __float_as_uint(__ldg(b.adv + off)));
    rec[3] = make_uint4(__float_as_uint(__ldg(b.ret + off)), 0u, 0u, 0u);
}

}  // namespace hfu

// once per rollout: record i of the (T*N) time-major samples
__global__ void rollout_pack_kernel(BatchDev b, uint4* __restrict__ packed) {
    const int64_t total = (int64_t)b.T * b.N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 rec[4];
        hfu::make_record(b, i, rec);
#pragma unroll
        for (int q = 0; q < 4; ++q) __stcs(packed + 4 * i + q, rec[q]);
    }
}
// callers without rollout records: the minibatch's own records, in minibatch order
__global__ void batch_pack_kernel(BatchDev b, const uint32_t* __restrict__ offs, uint4* __restrict__ packed) {
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= b.n) return;
    uint4 rec[4];
    hfu::make_record(b, (int64_t)offs[pos], rec);
#pragma unroll
    for (int q = 0; q < 4; ++q) packed[4 * pos + q] = rec[q];
}

// Development aid (build with GS_NVCC_EXTRA=-DGS_F16_TRACE): clock64() stamps of one tile of one CTA per set, compute warp 0 (role 0)
// and the MMA warp (role 1); read back with gs_debug_f16_trace().  Not part of the ABI; absent from normal builds.
#ifdef GS_F16_TRACE
#ifndef GS_F16_TRACE_TILE
#define GS_F16_TRACE_TILE 5
#endif
__device__ long long g_f16_trace[2][2][16];
__device__ long long g_f16_tiles[2][64];   // tile start stamps of compute warp 0 of each set (slot 62: kernel entry, 63: exit)
#define GS_TR(role, k) do { if (blockIdx.x == 7 && i == GS_F16_TRACE_TILE && lane == 0 && warp == ((role) == 1 ? C::kComputeWarps + set : set * C::kCW)) g_f16_trace[set][role][k] = clock64(); } while (0)
#else
#define GS_TR(role, k) do { } while (0)
#endif

template <int H, int ALGO, bool TRACK, int ACT>
__global__ void __launch_bounds__(hfu::Cfg<H>::kThreads, 1)
update_f16_kernel(MlpDev m, BatchDev b, HpDev hp, const double* __restrict__ adv_mom, const double* __restrict__ ret_mom,
                 const uint32_t* __restrict__ offs /* nullable: identity */, const uint4* __restrict__ records, float* __restrict__ grad_partials,
                 int64_t pstride, double* __restrict__ metric_partials, uint32_t* __restrict__ dead) {
    using C = hfu::Cfg<H>;
    using namespace hfu;
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(sm + C::oBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::oTmem);
    float* bhs = reinterpret_cast<float*>(sm + C::oBH);
    float* ncs = reinterpret_cast<float*>(sm + C::oNcs);
    double* red = reinterpret_cast<double*>(sm + C::oRed);
    float* fr = reinterpret_cast<float*>(sm + C::oFr);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = m.A, D = m.D;
#ifdef GS_F16_TRACE
    if (blockIdx.x == 7 && tid == 0) { g_f16_tiles[0][62] = clock64(); g_f16_tiles[1][62] = g_f16_tiles[0][62]; }
#endif

    // ---- prologue that does not depend on the previous kernel's results ----------------------------------------------------
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int s = 0; s < C::kSets; ++s)
            for (int k = 0; k < kBars; ++k)
                mbar_init(&bars_all[s * kBars + k], k < RDY_X ? 1u : ((k == RDY_X || k == RDY_G) ? 4u : (uint32_t)C::kCW));
        fence_mbar_init();
    }
    for (int s = 0; s < C::kSets; ++s) {
        uint4* x4 = reinterpret_cast<uint4*>(sm + s * C::kSetBytes + C::oX);
        for (int i = tid; i < (int)(kSlab / 16); i += C::kThreads) x4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < PM_N) red[tid] = 0.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");     // programmatic dependent launch: the weights below come from the previous step
    // ---- weights -> operand tiles ----------------------------------------------------------------------------------------------
    stage_weights<H>(m, sm + C::oW2hi, sm + C::oW2lo, sm + C::oWS, tid, C::kThreads);
    if (tid < 4) bhs[tid] = tid < A ? __ldg(m.bp + tid) : ((tid == A && m.has_value) ? __ldg(m.bv) : 0.f);
    if (tid == 0) {
        float adv_mean = 0.f, adv_den = 1.f, ret_mean = 0.f, ret_den = 1.f;
        if (hp.normalize_adv) norm_consts(adv_mom, adv_mean, adv_den);
        if (hp.normalize_ret) norm_consts(ret_mom, ret_mean, ret_den);
        ncs[0] = adv_mean; ncs[1] = adv_den; ncs[2] = ret_mean; ncs[3] = ret_den; ncs[4] = 1.0f / (float)b.n;
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(*tmem_slot);
    const uint32_t warp_u = uniform((uint32_t)warp);

    const int64_t n_tiles = (b.n + kRows - 1) / kRows;
    const int G = (int)gridDim.x;
    const int n_cta = (int)((n_tiles - blockIdx.x + G - 1) / G);                  // tiles of this CTA (>= 1)

    // ============================================ MMA-issuing warps ============================================================
    if (warp_u >= (uint32_t)C::kComputeWarps) {
        const int set = (int)warp_u - C::kComputeWarps;
        const int n_my = n_cta > set ? (n_cta - set + C::kSets - 1) / C::kSets : 0;
        uint64_t* bars = bars_all + set * kBars;
        const uint32_t S = smem_u32(sm) + (uint32_t)set * C::kSetBytes;
        const uint32_t sPhi = S + C::oPhi, sPlo = S + C::oPlo, sQhi = S + C::oQhi, sQlo = S + C::oQlo, sX = S + C::oX;
        const uint32_t sW2hi = smem_u32(sm) + C::oW2hi, sW2lo = smem_u32(sm) + C::oW2lo, sWS = smem_u32(sm) + C::oWS;
        const uint32_t T = tmem + (uint32_t)set * C::kSetCols;
        constexpr int KS = H / 16;                                                // k-steps over the feature dimension
        auto kfeat = [](int kk) -> uint32_t { return (uint32_t)(kk >> 2) * kSlab + (uint32_t)(kk & 3) * 32u; };          // activation tiles
        auto kw2 = [](int kk) -> uint32_t { return (uint32_t)(kk >> 2) * C::kW2Slab + (uint32_t)(kk & 3) * 32u; };       // W2 / WS rows = H
        auto issue_l1 = [&](int i) {                                              // z1(i) = x16(i) . [W1 | b1]: needs only the tile's record
            const uint32_t sXg = sX + 32u * (1u + ((uint32_t)i & 1u));
            if (elect_one()) {
                mma_f16(T + C::cAcc0, desc(sXg), desc(sWS + 32u), idesc_f16(128, H, 0, 0), 0u);      // small terms first: the accumulator
                mma_f16(T + C::cAcc0, desc(sXg), desc(sWS), idesc_f16(128, H, 0, 0), 1u);            // truncates relative to its magnitude
                mma_commit(&bars[BAR_Z1]);
            }
            __syncwarp();
        };
        if (n_my > 0) {
            mbar_wait(&bars[RDY_X], 0); fence_after_sync();
            issue_l1(0);
        }
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const uint32_t p = (uint32_t)i & 1u;
            const uint32_t sXg = sX + 32u * (1u + p);                             // x16 group of this tile
            const uint32_t acc_w = (i % kFlushTiles) != 0 ? 1u : 0u;              // weight-gradient accumulators restart after a flush
            GS_TR(1, 1);
            mbar_wait(&bars[RDY_H1], p); fence_after_sync();
            GS_TR(1, 2);
            if (elect_one()) {                                                    // bias + fwd
                mma_f16(T + C::cAcc1, desc(sXg), desc(sWS + 64u), idesc_f16(128, H, 0, 0), 0u);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a0 = pass == 0 ? sPlo : sPhi, b0 = pass == 1 ? sW2lo : sW2hi;      // lo*hi, hi*lo, hi*hi
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) mma_f16(T + C::cAcc1, desc(a0 + kfeat(kk)), desc(b0 + kw2(kk)), idesc_f16(128, H, 0, 0), 1u);
                }
                mma_commit(&bars[BAR_Z2]);
            }
            __syncwarp();
            GS_TR(1, 3);
            mbar_wait(&bars[RDY_H2], p); fence_after_sync();
            GS_TR(1, 4);
            if (elect_one()) {                                                    // heads: B = WS group 3 read MN-major (N = 16, K = feature rows)
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const uint32_t a0 = pass == 0 ? sQlo : sQhi;
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk)
                        mma_f16(T + C::cH, desc(a0 + kfeat(kk)), desc(sWS + 96u + (uint32_t)kk * 2048u, C::kW2Slab), idesc_f16(128, 16, 0, 1), (pass | kk) ? 1u : 0u);
                }
                mma_commit(&bars[BAR_OUT]);
            }
            __syncwarp();
            GS_TR(1, 5);
            mbar_wait(&bars[RDY_G], p); fence_after_sync();
            GS_TR(1, 6);
            if (elect_one()) {
                mma_f16(T + C::cAcc0, desc(sX), desc(sWS + 96u), idesc_f16(128, H, 0, 0), 0u);                   // dh2 = g16 . WS3
                mma_commit(&bars[BAR_DH2]);
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {                            // W-c: dWh^T += h2^T g16
                    const uint32_t a0 = pass == 0 ? sQlo : sQhi;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        mma_f16(T + C::cWh, desc(a0 + (uint32_t)kk * 2048u, kSlab), desc(sX + (uint32_t)kk * 2048u, kSlab), idesc_f16(H, 16, 1, 1),
                                 (pass | kk) ? 1u : acc_w);
                }
                mma_commit(&bars[BAR_WC]);
            }
            __syncwarp();
            GS_TR(1, 7);
            mbar_wait(&bars[RDY_DZ2], p); fence_after_sync();
            GS_TR(1, 8);
            if (elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {                            // dgrad: B = W2 read MN-major (N = k, K = j rows)
                    const uint32_t a0 = pass == 0 ? sQlo : sQhi, b0 = pass == 1 ? sW2lo : sW2hi;
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk)
                        mma_f16(T + C::cAcc1, desc(a0 + kfeat(kk)), desc(b0 + (uint32_t)kk * 2048u, C::kW2Slab), idesc_f16(128, H, 0, 1), (pass | kk) ? 1u : 0u);
                }
                mma_commit(&bars[BAR_DH1]);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {                            // W-a: [dW2 | db2] += dz2^T [h1 | g16]
                    const uint32_t a0 = pass == 0 ? sQlo : sQhi, b0 = pass == 1 ? sPlo : sPhi;        // the first pass covers all H + 16 columns
                    const uint32_t id = pass == 1 ? idesc_f16(H, H, 1, 1) : idesc_f16(H, H + 16, 1, 1);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        mma_f16(T + C::cW2, desc(a0 + (uint32_t)kk * 2048u, kSlab), desc(b0 + (uint32_t)kk * 2048u, kSlab), id, (pass | kk) ? 1u : acc_w);
                }
                mma_commit(&bars[BAR_WA]);
            }
            __syncwarp();
            if (i + 1 < n_my) {                                                   // the next tile's layer 1 goes ahead of this tile's last group:
                mbar_wait(&bars[RDY_X], p ^ 1u); fence_after_sync();              // acc0 is free (dh2 was read in stage D), its record has landed
                issue_l1(i + 1);
            }
            GS_TR(1, 9);
            mbar_wait(&bars[RDY_DZ1], p); fence_after_sync();
            GS_TR(1, 10);
            if (elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {                            // W-b: [dW1 | db1] += dz1^T x16   (dz1 lives in Q)
                    const uint32_t a0 = pass == 0 ? sQlo : sQhi;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        mma_f16(T + C::cW1, desc(a0 + (uint32_t)kk * 2048u, kSlab), desc(sXg + (uint32_t)kk * 2048u, kSlab), idesc_f16(H, 16, 1, 1),
                                 (pass | kk) ? 1u : acc_w);
                }
                mma_commit(&bars[BAR_WB]);
            }
            __syncwarp();
            GS_TR(1, 11);
        }
        return;
    }

    // ============================================ compute warps ==================================================================
    const int set = warp / C::kCW, ws = warp % C::kCW;
    const int quad = ws & 3, cg = ws >> 2;                    // TMEM lane quadrant (== warp & 3), 32-column group
    const int row = quad * 32 + lane;                         // sample row of the tile == TMEM lane
    const bool loss_thread = cg == 0;
    const int n_my = n_cta > set ? (n_cta - set + C::kSets - 1) / C::kSets : 0;
    uint64_t* bars = bars_all + set * kBars;
    unsigned char* S = sm + (size_t)set * C::kSetBytes;
    unsigned char* Xrow = S + C::oX + row * 128;
    const int sw = row & 7;
    // this thread's four 16-byte chunks (32 columns) of its row in an activation tile
    const uint32_t my_off = (uint32_t)(cg >> 1) * kSlab + (uint32_t)row * 128u;
    const int c0 = (cg & 1) * 4;
    const uint32_t T = tmem + (uint32_t)set * C::kSetCols + ((uint32_t)(quad * 32) << 16);
    float* out = grad_partials + (size_t)(blockIdx.x * C::kSets + set) * pstride;   // this set's partial gradient vector
    const ParamOffsets po = param_offsets(D, H, H, A, m.has_value);

    float pm[PM_N];
#pragma unroll
    for (int q = 0; q < PM_N; ++q) pm[q] = 0.f;
    float gsum[kNH] = {0.f, 0.f, 0.f, 0.f};
    float zs0 = 0.f, zq0 = 0.f, zs1 = 0.f, zq1 = 0.f;
    uint32_t dead0 = 0, dead1 = 0;                            // lane i: dead-sample count of neuron 32 cg + i

    auto tile_of = [&](int i) -> int64_t { return (int64_t)blockIdx.x + (int64_t)(C::kSets * i + set) * G; };
    // cp.async of a sample record into the X row: x16 -> group 1 + (i & 1), scalars -> group 3
    auto prefetch = [&](int i, uint32_t off, bool ok) {
        const int gx = 2 * (1 + (i & 1));
        const uint32_t d0 = smem_u32(Xrow + (((gx) ^ sw) << 4)), d1 = smem_u32(Xrow + (((gx + 1) ^ sw) << 4));
        if (ok) {
            const uint4* src = records + (size_t)off * 4;
            cp_async16(d0, src); cp_async16(d1, src + 1);
            cp_async16(smem_u32(Xrow + ((6 ^ sw) << 4)), src + 2); cp_async16(smem_u32(Xrow + ((7 ^ sw) << 4)), src + 3);
        } else {                                              // rows past the minibatch: all-zero x16 (no ones) -> every activation and gradient is 0
            *reinterpret_cast<uint4*>(Xrow + ((gx ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(Xrow + (((gx + 1) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
    };
    auto offset_of = [&](int i, bool& ok) -> uint32_t {
        const int64_t pos = tile_of(i) * kRows + row;
        ok = i < n_my && pos < b.n;
        return ok ? (offs ? __ldg(offs + pos) : (uint32_t)pos) : 0u;
    };
    // fold the set's weight-gradient accumulators into its partial vector (first: overwrite)
    auto flush = [&](bool first) {
        const float invB = ncs[4];
        const bool owner = (H == 128) || lane < 16;           // M = 64 accumulators: row m in lane (m / 16) * 32 + m % 16
        const int mrow = H == 128 ? row : quad * 16 + (lane & 15);
        {
            float v[32];
            tmem_ld32(T + C::cW2 + 32 * cg, v);
            float* d1 = out + po.w2 + (int64_t)mrow * H + 32 * cg;
            const bool vec = (po.w2 & 3) == 0;
            float4 o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (owner && !first) {                            // every old value is requested before the first use: one L2 round trip
                if (vec) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = __ldcg(reinterpret_cast<const float4*>(d1) + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = make_float4(__ldcg(d1 + 4 * q), __ldcg(d1 + 4 * q + 1), __ldcg(d1 + 4 * q + 2), __ldcg(d1 + 4 * q + 3));
                }
            }
            tmem_ld_wait();
            if (owner) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    o[q].x = fmaf(v[4 * q], invB, o[q].x); o[q].y = fmaf(v[4 * q + 1], invB, o[q].y);
                    o[q].z = fmaf(v[4 * q + 2], invB, o[q].z); o[q].w = fmaf(v[4 * q + 3], invB, o[q].w);
                    if (vec) reinterpret_cast<float4*>(d1)[q] = o[q];
                    else { d1[4 * q] = o[q].x; d1[4 * q + 1] = o[q].y; d1[4 * q + 2] = o[q].z; d1[4 * q + 3] = o[q].w; }
                }
            }
        }
        if (cg == 0) {
            float e[16], w1[16], wh[16];
            tmem_ld16(T + C::cW2 + H, e);
            tmem_ld16(T + C::cW1, w1);
            tmem_ld16(T + C::cWh, wh);
            // destinations of this row's 13 small entries: b2, w1[0..6], b1, wp[0..2], wv (unused ones point at b2: weight 0)
            int64_t idx[13];
            idx[0] = po.b2 + mrow;
#pragma unroll
            for (int d = 0; d < kMaxD; ++d) idx[1 + d] = d < D ? po.w1 + (int64_t)mrow * D + d : idx[0];
            idx[8] = po.b1 + mrow;
#pragma unroll
            for (int r = 0; r < 3; ++r) idx[9 + r] = r < A ? po.wp + (int64_t)r * H + mrow : idx[0];
            idx[12] = m.has_value ? po.wv + mrow : idx[0];
            float old[13];
#pragma unroll
            for (int q = 0; q < 13; ++q) old[q] = (owner && !first) ? __ldcg(out + idx[q]) : 0.f;
            tmem_ld_wait();
            if (owner) {
                float val[13];
                val[0] = e[14];
#pragma unroll
                for (int d = 0; d < kMaxD; ++d) val[1 + d] = w1[d] + w1[7 + d];
                val[8] = w1[14];
#pragma unroll
                for (int r = 0; r < 3; ++r) val[9 + r] = wh[r] + wh[4 + r];
                val[12] = A == 2 ? wh[2] + wh[6] : wh[3] + wh[7];
#pragma unroll
                for (int q = 12; q >= 0; --q) {               // q = 0 (b2) last: the unused slots alias it
                    const bool used = q == 0 || q == 8 || (q >= 1 && q <= 7 && q - 1 < D) || (q >= 9 && q <= 11 && q - 9 < A) || (q == 12 && m.has_value);
                    if (used) out[idx[q]] = fmaf(val[q], invB, old[q]);
                }
            }
        }
        fence_before_sync();
    };
    // stage closer: this warp's operand stores / TMEM reads are done
    auto warp_ready = [&](int which) {
        fence_proxy_async();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[which]);
    };
    // forward stage: 32 pre-activations of the row -> activation -> (hi, lo) -> tile
    auto fwd_stage = [&](uint32_t acc_col, unsigned char* t_hi, unsigned char* t_lo, bool valid, float& zs, float& zq, uint32_t& dcnt,
                         uint64_t* release, uint32_t parity) {
        float z[32];
        tmem_ld32(T + acc_col + 32 * cg, z);
        tmem_ld_wait();
        if (TRACK) {
            float mn = 1.0f;
            if (valid) {
#pragma unroll
                for (int q = 0; q < 32; ++q) { zs += z[q]; zq = fmaf(z[q], z[q], zq); mn = fminf(mn, fabsf(z[q])); }
            }
            if (__any_sync(0xffffffffu, mn < 1e-6f)) {
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const uint32_t hits = __popc(__ballot_sync(0xffffffffu, valid && fabsf(z[q]) < 1e-6f));
                    if (lane == q) dcnt += hits;
                }
            }
        }
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) split_pair(act_fwd(z[2 * e], ACT), act_fwd(z[2 * e + 1], ACT), hw[e], lw[e]);
        if (release) mbar_wait(release, parity);              // the MMAs still reading the tile's previous content
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t o = my_off + (uint32_t)(((c0 + c) ^ sw) << 4);
            *reinterpret_cast<uint4*>(t_hi + o) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
            *reinterpret_cast<uint4*>(t_lo + o) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
        }
    };
    // backward stage: d(loss)/d(activation) of the row * act'(activation held in the tile) -> (hi, lo) -> the same tile, once
    // `release` says the MMAs still reading the tile are done
    auto bwd_stage = [&](uint32_t acc_col, const unsigned char* t_hi, const unsigned char* t_lo, unsigned char* d_hi, unsigned char* d_lo,
                         uint64_t* release, uint32_t parity) {
        float d[32];
        tmem_ld32(T + acc_col + 32 * cg, d);
        uint4 hq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) hq[c] = *reinterpret_cast<const uint4*>(t_hi + my_off + (uint32_t)(((c0 + c) ^ sw) << 4));
        tmem_ld_wait();
        if (ACT == GS_ACT_RELU) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t w[4] = {hq[c].x, hq[c].y, hq[c].z, hq[c].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    d[8 * c + 2 * e] = (w[e] & 0x7FFFu) ? d[8 * c + 2 * e] : 0.f;
                    d[8 * c + 2 * e + 1] = (w[e] & 0x7FFF0000u) ? d[8 * c + 2 * e + 1] : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 lq = *reinterpret_cast<const uint4*>(t_lo + my_off + (uint32_t)(((c0 + c) ^ sw) << 4));
                const uint32_t w[4] = {hq[c].x, hq[c].y, hq[c].z, hq[c].w}, l[4] = {lq.x, lq.y, lq.z, lq.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float h0, h1, l0, l1;
                    unpack_pair(w[e], h0, h1);
                    unpack_pair(l[e], l0, l1);
                    h0 += l0; h1 += l1;
                    d[8 * c + 2 * e] *= 1.0f - h0 * h0;
                    d[8 * c + 2 * e + 1] *= 1.0f - h1 * h1;
                }
            }
        }
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) split_pair(d[2 * e], d[2 * e + 1], hw[e], lw[e]);
        mbar_wait(release, parity);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t o = my_off + (uint32_t)(((c0 + c) ^ sw) << 4);
            *reinterpret_cast<uint4*>(d_hi + o) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
            *reinterpret_cast<uint4*>(d_lo + o) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
        }
    };

    uint32_t next_off = 0;
    bool next_ok = false;
    if (n_my > 0 && loss_thread) {
        bool ok;
        const uint32_t off = offset_of(0, ok);
        prefetch(0, off, ok);
        next_off = offset_of(1, next_ok);
        cp_async_wait_all();
        fence_proxy_async();
        // Two sets that start together stay in lockstep (both in a SIMT stage, then both queueing MMAs: measured, clock64 trace): the
        // second set starts half a tile late -- when the first set's head outputs are done -- so one set's MMA groups run under the
        // other's SIMT stages, and that offset is self-sustaining.
        if (C::kSets == 2 && set == 1) mbar_wait(&bars_all[BAR_OUT], 0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[RDY_X]);
    }
    if (n_my == 0) {                                          // a set without tiles still owns a partial vector: zeros
        for (int64_t i = ws * 32 + lane; i < po.total; i += C::kCW * 32) out[i] = 0.f;
    }
#pragma unroll 1
    for (int i = 0; i < n_my; ++i) {
        const uint32_t p = (uint32_t)i & 1u;
        const bool valid = tile_of(i) * kRows + row < b.n;
        // ---- A: h1 ----------------------------------------------------------------------------------------------------------
#ifdef GS_F16_TRACE
        if (blockIdx.x == 7 && lane == 0 && ws == 0 && i < 62) g_f16_tiles[set][i] = clock64();
#endif
        GS_TR(0, 0);
        mbar_wait(&bars[BAR_Z1], p);
        fence_after_sync();
        GS_TR(0, 1);
        if (i > 0 && (i % kFlushTiles) == 0) {               // every weight-gradient MMA of the previous tiles has completed
            mbar_wait(&bars[BAR_WB], p ^ 1u);
            fence_after_sync();
            flush(i == kFlushTiles);
        }
        fwd_stage(C::cAcc0, S + C::oPhi, S + C::oPlo, valid, zs0, zq0, dead0, nullptr, 0u);   // P is free: W-a(i-1) was awaited in stage E
        warp_ready(RDY_H1);
        // ---- B: h2 ----------------------------------------------------------------------------------------------------------
        GS_TR(0, 2);
        mbar_wait(&bars[BAR_Z2], p);
        fence_after_sync();
        GS_TR(0, 3);
        fwd_stage(C::cAcc1, S + C::oQhi, S + C::oQlo, valid, zs1, zq1, dead1, i > 0 ? &bars[BAR_WB] : nullptr, p ^ 1u);   // W-b(i-1) reads dz1 from Q
        warp_ready(RDY_H2);
        // ---- C: loss (one thread per row) ----------------------------------------------------------------------------------------
        GS_TR(0, 4);
        if (loss_thread) {
            mbar_wait(&bars[BAR_OUT], p);
            fence_after_sync();
            GS_TR(0, 5);
            float c[16];
            tmem_ld16(T + C::cH, c);
            const uint4 sc = *reinterpret_cast<const uint4*>(Xrow + ((6 ^ sw) << 4));
            const float ret_s = *reinterpret_cast<const float*>(Xrow + ((7 ^ sw) << 4));
            tmem_ld_wait();
            GS_TR(0, 11);
            float g[kNH] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                float outv[kNH];
#pragma unroll
                for (int r = 0; r < kNH; ++r) outv[r] = bhs[r] + (c[r] + c[8 + r]);
                sample_loss<ALGO>(outv, A, (int)sc.x, __uint_as_float(sc.y), __uint_as_float(sc.z), __uint_as_float(sc.w), ret_s, hp, ncs[0], ncs[1],
                                  ncs[2], ncs[3], 1.0f /* sum loss: the 1/B is applied when the accumulators are folded */, g, pm);
#pragma unroll
                for (int r = 0; r < kNH; ++r) gsum[r] += g[r];
            }
            GS_TR(0, 12);
            uint32_t h01, l01, h23, l23;
            split_pair(g[0], g[1], h01, l01);
            split_pair(g[2], g[3], h23, l23);
            *reinterpret_cast<uint4*>(Xrow + ((0 ^ sw) << 4)) = make_uint4(h01, h23, l01, l23);
            *reinterpret_cast<uint4*>(Xrow + ((1 ^ sw) << 4)) = make_uint4(h01, h23, 0u, kOnes2);
            GS_TR(0, 13);
            if (i + 1 < n_my) prefetch(i + 1, next_off, next_ok);   // the scalars above were consumed: their slot may be overwritten
            GS_TR(0, 14);
            next_off = offset_of(i + 2, next_ok);
            warp_ready(RDY_G);
            GS_TR(0, 15);
        }
        // ---- D: dz2 ---------------------------------------------------------------------------------------------------------
        GS_TR(0, 6);
        mbar_wait(&bars[BAR_DH2], p);
        fence_after_sync();
        GS_TR(0, 7);
        bwd_stage(C::cAcc0, S + C::oQhi, S + C::oQlo, S + C::oQhi, S + C::oQlo, &bars[BAR_WC], p);
        warp_ready(RDY_DZ2);
        // ---- E: dz1 ---------------------------------------------------------------------------------------------------------
        if (loss_thread && i + 1 < n_my) {                    // the next tile's record (cp.async issued in stage C) has landed: its layer 1 may start
            cp_async_wait_all();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[RDY_X]);
        }
        GS_TR(0, 8);
        mbar_wait(&bars[BAR_DH1], p);
        fence_after_sync();
        GS_TR(0, 9);
        bwd_stage(C::cAcc1, S + C::oPhi, S + C::oPlo, S + C::oQhi, S + C::oQlo, &bars[BAR_WA], p);   // act'(h1) from P; dz1 -> Q once W-a is done with dz2
        warp_ready(RDY_DZ1);
        GS_TR(0, 10);
    }
#ifdef GS_F16_TRACE
    if (blockIdx.x == 7 && lane == 0 && ws == 0) g_f16_tiles[set][61] = clock64();
#endif
    if (n_my > 0) {
        mbar_wait(&bars[BAR_WB], (uint32_t)(n_my - 1) & 1u);
        fence_after_sync();
        flush(n_my <= kFlushTiles);
    }
    // ---- CTA reductions: metric partials (doubles, shared-memory atomics), activation statistics, head-bias gradients --------
    if (loss_thread) {
#pragma unroll
        for (int q = 0; q < PM_N; ++q) {
            const float v = warp_sum(pm[q]);
            if (lane == 0 && v != 0.f) atomicAdd(red + q, (double)v);
        }
#pragma unroll
        for (int r = 0; r < kNH; ++r) {
            const float v = warp_sum(gsum[r]);
            if (lane == 0) fr[(set * 4 + quad) * 4 + r] = v;
        }
    }
    if (TRACK) {
        const double a0 = warp_sum((double)zs0), a1 = warp_sum((double)zq0), a2 = warp_sum((double)zs1), a3 = warp_sum((double)zq1);
        if (lane == 0) { atomicAdd(red + PM_Z0, a0); atomicAdd(red + PM_Z0SQ, a1); atomicAdd(red + PM_Z1, a2); atomicAdd(red + PM_Z1SQ, a3); }
        if (dead0) atomicAdd(dead + 32 * cg + lane, dead0);
        if (dead1) atomicAdd(dead + H + 32 * cg + lane, dead1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(C::kCompute) : "memory");
    if (tid < PM_N) metric_partials[(size_t)blockIdx.x * PM_N + tid] = red[tid];
    if (ws == 0 && lane < kNH) {                              // fixed order: deterministic
        const float sgm = (fr[(set * 4 + 0) * 4 + lane] + fr[(set * 4 + 1) * 4 + lane]) + (fr[(set * 4 + 2) * 4 + lane] + fr[(set * 4 + 3) * 4 + lane]);
        if (lane < A) out[po.bp + lane] = sgm * ncs[4];
        else if (lane == A && m.has_value) out[po.bv] = sgm * ncs[4];
    }
#ifdef GS_F16_TRACE
    if (blockIdx.x == 7 && lane == 0 && ws == 0) g_f16_tiles[set][63] = clock64();
#endif
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host launchers (called from update_kernels.cu) -----------------------------------------------------------------------------
int f16_sets(int H) { return H == 64 ? 2 : 1; }   // partial gradient vectors per CTA (update_wide.cu's 256-wide kernels: 1)

template <int H, int ALGO>
static int launch_f16_h(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                       const uint32_t* offs, const void* records, float* grad_partials, int64_t pstride, double* metric_partials, uint32_t* dead,
                       int grid, cudaStream_t st) {
    using C = hfu::Cfg<H>;
    using KernelFn = void (*)(MlpDev, BatchDev, HpDev, const double*, const double*, const uint32_t*, const uint4*, float*, int64_t, double*, uint32_t*);
    KernelFn kern;
    if (md.act == GS_ACT_RELU) kern = track ? update_f16_kernel<H, ALGO, true, GS_ACT_RELU> : update_f16_kernel<H, ALGO, false, GS_ACT_RELU>;
    else kern = track ? update_f16_kernel<H, ALGO, true, GS_ACT_TANH> : update_f16_kernel<H, ALGO, false, GS_ACT_TANH>;
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    // Programmatic dependent launch: the CTAs may start while the previous kernel of the stream (the step tail of the previous
    // minibatch, which triggers early) is still in its serial phase; everything up to griddepcontrol.wait -- TMEM allocation,
    // barrier init, zeroing -- runs under it, the weights are read after it.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)C::kThreads); cfg.dynamicSmemBytes = C::kSmemBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GS_CUDA(cudaLaunchKernelEx(&cfg, kern, md, b, hp, adv_mom, ret_mom, offs, reinterpret_cast<const uint4*>(records), grad_partials, pstride,
                               metric_partials, dead));
    return 0;
}

template <int ALGO>
int launch_update_f16(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                     const uint32_t* offs, const void* records, float* grad_partials, int64_t pstride, double* metric_partials, uint32_t* dead,
                     int grid, cudaStream_t st) {
    if (md.H1 == 64) return launch_f16_h<64, ALGO>(md, b, hp, track, adv_mom, ret_mom, offs, records, grad_partials, pstride, metric_partials, dead, grid, st);
    return launch_f16_h<128, ALGO>(md, b, hp, track, adv_mom, ret_mom, offs, records, grad_partials, pstride, metric_partials, dead, grid, st);
}
template int launch_update_f16<ALGO_PPO>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, const uint32_t*, const void*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);
template int launch_update_f16<ALGO_REINFORCE>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, const uint32_t*, const void*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);

int launch_rollout_pack(const BatchDev& b, void* packed, int device, cudaStream_t st) {
    const int64_t total = (int64_t)b.T * b.N;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = 16ll * sm_count(device);
    if (blocks > cap) blocks = cap;
    rollout_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, reinterpret_cast<uint4*>(packed));
    GS_LAUNCH_CHECK();
    return 0;
}
int launch_batch_pack(const BatchDev& b, const uint32_t* offs, void* packed, cudaStream_t st) {
    batch_pack_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, st>>>(b, offs, reinterpret_cast<uint4*>(packed));
    GS_LAUNCH_CHECK();
    return 0;
}

}  // namespace gs

#ifdef GS_F16_TRACE
extern "C" int gs_debug_f16_trace(long long* host_out /* [2][2][16] */) {
    return (int)cudaMemcpyFromSymbol(host_out, gs::g_f16_trace, sizeof(long long) * 64);
}
extern "C" int gs_debug_f16_tiles(long long* host_out /* [2][64] */) {
    return (int)cudaMemcpyFromSymbol(host_out, gs::g_f16_tiles, sizeof(long long) * 128);
}
#endif
