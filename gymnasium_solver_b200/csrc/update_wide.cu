// update_wide.cu — tcgen05 / TMEM minibatch update (PPO and REINFORCE) for the 256 x 256 relu MLP (`mlp_medium`: the reference's
// shipped CartPole-v1 and MountainCar-v0 configurations, utils/model_registry.py:28-31).
//
// Same contract and the same fp16x3 arithmetic as update_f16.cu (hi = fp16(x), lo = fp16(x - hi); lo*hi + hi*lo + hi*hi with fp32
// accumulation in TMEM; backward quantities carried for the SUM loss, 1/B applied when accumulators are folded; reference:
// agents/ppo/ppo_agent.py:21-152, agents/reinforce/reinforce_agent.py:11-88).  What changes at H = 256 is where things fit:
//   * W2 as operand tiles (hi + lo) is 256 KB -- more than an SM's shared memory.  `stage_w2_kernel` writes W2 (rows j) and W2^T (rows k)
//     ONCE per step into the workspace in the operand layout ([precision][64-column slab][256 rows][128 B, SWIZZLE_128B chunks]) and every
//     CTA streams them through a 3 x 16 KB ring with `cp.async.bulk` (a loader warp, full / empty mbarriers).  Both GEMM phases read B
//     MN-major (K = row) with N = 256, so one ring stage is 32 K-rows of all four slabs (two K = 16 steps): 16 stages per tile for the
//     forward pass (W2^T: lo blocks then hi blocks) and 16 for dgrad (W2).  CTAs walk the 16 blocks of a phase from rotated starts.
//   * the layer-2 weight gradient needs a 256 x 256 fp32 accumulator = all 512 TMEM columns.  It gets its own kernel: the fused
//     kernel stores the two operands of that product -- the h1 and dz2 tiles exactly as they sit in shared memory (256 KB per
//     128-sample tile, `cp.async.bulk` shared -> global issued by the MMA warp in 16 KB pieces between the stages of the GEMM phase that
//     follows) -- and `wgrad_wide_kernel` streams them back (3-stage ring of 32-sample slices, both operands MN-major straight from the
//     stored bytes) into dW2 = dz2^T . h1 with M = 2 x 128, N = 256.  Every other gradient (dW1, b1, b2, heads) is a 16-column
//     accumulator of the fused kernel, as in update_f16.cu; b2's is dz2^T . [ones] (the ones of the g16 group).
//   * one activation buffer P ([128][256] hi + lo = 128 KB) holds h1, then h2, then dz2, then dz1; relu' of both layers survives as a
//     64-bit mask per thread.  One 256-column accumulator serves z1, z2, dh2, dh1 in turn, so the MMA phases and the SIMT stages of a
//     tile alternate.  The hi half of the tile is also kept in tensor memory (two fp16 per column) as the A operand of the K-major MMAs.
// Thread = (sample row, 64-column slab), walking its slab 16 columns at a time; 16 compute warps + MMA warp + loader warp, one CTA per SM,
// persistent over the tiles.  Measurements and the steps that led here: DESIGN.md section 3, profiles/r2_update_wide_ncu_details.md.
#include <type_traits>

#define GS_FAST_TRANSCENDENTALS
#include "wide.cuh"

namespace gs {

using namespace tc;

// W2 (rows j, columns k: dgrad's B) and W2^T (rows k, columns j: the forward pass's B) -> operand layout in the workspace, each as
// hi then lo, once per step.  Both are read MN-major (K = row), so one 16 KB ring stage = 32 rows of all four 64-column slabs.
__global__ void stage_w2_kernel(MlpDev m, unsigned char* __restrict__ staged) {
    using namespace wfu;
    const int n_threads = (int)(gridDim.x * blockDim.x);
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < H * H / 4; i += n_threads) {
        const int j = i / (H / 4), k = 4 * (i % (H / 4));
        const float4 w = __ldg(reinterpret_cast<const float4*>(m.w2 + j * H) + (i % (H / 4)));
        uint32_t h0, l0, h1, l1;
        hfu::split_pair(w.x, w.y, h0, l0);
        hfu::split_pair(w.z, w.w, h1, l1);
        const uint32_t o = hfu::tile_off(j, k, H);
        *reinterpret_cast<uint2*>(staged + o) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(staged + kW2Prec + o) = make_uint2(l0, l1);
        const uint32_t hv[4] = {h0 & 0xFFFFu, h0 >> 16, h1 & 0xFFFFu, h1 >> 16}, lv[4] = {l0 & 0xFFFFu, l0 >> 16, l1 & 0xFFFFu, l1 >> 16};
#pragma unroll
        for (int e = 0; e < 4; ++e) {                          // transposed copy: element (k + e, j)
            const uint32_t ot = hfu::tile_off(k + e, j, H);
            *reinterpret_cast<unsigned short*>(staged + 2 * kW2Prec + ot) = (unsigned short)hv[e];
            *reinterpret_cast<unsigned short*>(staged + 3 * kW2Prec + ot) = (unsigned short)lv[e];
        }
    }
}

// Development aid (GS_NVCC_EXTRA=-DGS_WIDE_TRACE): clock64() stamps of one tile of CTA 7 -- compute warp 0 (role 0) and the MMA warp (role 1)
#ifdef GS_WIDE_TRACE
__device__ long long g_wide_trace[2][24];
#define GS_WT(role, k) do { if (blockIdx.x == 7 && i == 5 && lane == 0 && (role == 1 || warp == 0)) g_wide_trace[role][k] = clock64(); } while (0)
#else
#define GS_WT(role, k) do { } while (0)
#endif

template <int ALGO, bool TRACK>
__global__ void __launch_bounds__(wfu::kWideThreads, 1)
update_wide_kernel(MlpDev m, BatchDev b, HpDev hp, const double* __restrict__ adv_mom, const double* __restrict__ ret_mom,
                   const uint32_t* __restrict__ offs /* nullable: identity */, const uint4* __restrict__ records,
                   const unsigned char* __restrict__ w2s, unsigned char* __restrict__ tiles, float* __restrict__ grad_partials, int64_t pstride,
                   double* __restrict__ metric_partials, uint32_t* __restrict__ dead) {
    using namespace wfu;
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + oTmem);
    float* bhs = reinterpret_cast<float*>(sm + oBH);
    float* ncs = reinterpret_cast<float*>(sm + oNcs);
    double* red = reinterpret_cast<double*>(sm + oRed);
    float* fr = reinterpret_cast<float*>(sm + oFr);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = m.A, D = m.D;

    // ---- prologue that does not depend on the previous kernel's results ----------------------------------------------------
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int k = 0; k < kBars; ++k)
            mbar_init(&bars[k], (k < RDY_X || k >= BAR_FULL) ? 1u : ((k == RDY_X || k == RDY_G) ? 4u : (uint32_t)kCW));
        fence_mbar_init();
    }
    {
        uint4* x4 = reinterpret_cast<uint4*>(sm + oX);
        for (int i = tid; i < (int)(kSlab / 16); i += kWideThreads) x4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < PM_N) red[tid] = 0.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");     // the weights (and the staged W2) come from the kernels before this one
    stage_ws<H>(m, sm + oWS, tid);
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
    const int n_my = (int)((n_tiles - blockIdx.x + G - 1) / G);                   // tiles of this CTA (>= 1)
    const uint32_t sBase = smem_u32(sm);
    const uint32_t sPhi = sBase + oPhi, sPlo = sBase + oPlo, sX = sBase + oX, sWS = sBase + oWS, sRing = sBase + oRing;
    const int rot = (int)(blockIdx.x & 15u);                                      // block order rotation of this CTA (loader and MMA warp agree)

    // ============================================ loader warp: W2 blocks -> ring ===============================================
    // stage = K rows 32 blk .. 32 blk + 31 of one precision of W2^T (forward: rows k) or W2 (dgrad: rows j), all four 64-column slabs
    if (warp_u == (uint32_t)kCW + 1u) {
        if (lane == 0) {
            uint32_t q = 0;
#pragma unroll 1
            for (int i = 0; i < n_my; ++i) {
#pragma unroll 1
                for (int s = 0; s < kStagesPerTile; ++s, ++q) {
                    const uint32_t slot = q % kRing, round = q / kRing;
                    mbar_wait(&bars[BAR_EMPTY + slot], (round & 1u) ^ 1u);
#ifdef GS_WIDE_NOLOAD   // timing experiment: the ring is only filled for the first tile (results are wrong)
                    if (i > 0) { mbar_arrive(&bars[BAR_FULL + slot]); continue; }
#endif
                    mbar_expect_tx(&bars[BAR_FULL + slot], kStage);
                    const int s16 = (s + rot) & 15;                               // CTAs walk the 16 blocks of a phase from different starts
                    const uint32_t prec = s16 < 8 ? 1u : 0u, blk = (uint32_t)s16 & 7u;
                    const unsigned char* src = w2s + (s < 16 ? 2u * kW2Prec : 0u) + prec * kW2Prec + blk * 4096u;
                    const uint32_t dst = sRing + slot * kStage;
#pragma unroll
                    for (uint32_t sl = 0; sl < 4u; ++sl) bulk_g2s(dst + sl * 4096u, src + sl * (H * 128u), 4096u, &bars[BAR_FULL + slot]);
                }
            }
        }
        return;
    }

    // ============================================ MMA-issuing warp ==============================================================
    // Warp-uniform control flow; every MMA group is issued by the elected lane under elect.sync (one UTCHMMA per instruction instead of the
    // per-instruction election loop a thread-dependent branch compiles to).  The bulk stores of the tiles are lane 0's (bulk groups are
    // per thread).  The tensor pipe executes a CTA's MMAs in issue order, so a commit also stands for the groups issued before it.
    if (warp_u == (uint32_t)kCW) {
        const uint32_t T = tmem;
        uint32_t q = 0;                                                           // ring stage counter (same sequence as the loader's)
        // descriptors: the start-address field counts 16-byte units, so operands are stepped by adding (bytes >> 4)
        const uint64_t dPhi = desc(sPhi), dPlo = desc(sPlo);                      // K-major A (rows = samples)
        const uint64_t dPhiT = desc(sPhi, kSlab), dPloT = desc(sPlo, kSlab);      // MN-major A (M = features: LBO = slab stride)
        const uint64_t dG = desc(sX, kSlab);                                      // g16 group as MN-major B of W-c / W-d
        const uint64_t dWS = desc(sWS), dWSh = desc(sWS + 96u, H * 128u);         // WS groups; head rows as MN-major B
        constexpr uint32_t idFull = idesc_f16(128, H, 0, 0), idFullT = idesc_f16(128, H, 0, 1);
        constexpr uint32_t idHeads = idesc_f16(128, 16, 0, 1), idW = idesc_f16(128, 16, 1, 1);
        auto kfeat = [](int kk) -> uint64_t { return (uint64_t)(((uint32_t)(kk >> 2) * kSlab + (uint32_t)(kk & 3) * 32u) >> 4); };
        auto issue_l1 = [&](int i) {
            const uint64_t dXg = desc(sX + 32u * (1u + ((uint32_t)i & 1u)));
            if (elect_one()) {
                mma_f16(T + cAcc, dXg, dWS + 2u, idFull, 0u);                     // x_hi . W1_lo first (small terms first)
                mma_f16(T + cAcc, dXg, dWS, idFull, 1u);
                mma_commit(&bars[BAR_Z1]);
            }
            __syncwarp();
        };
        // 32 small MMAs of a 16-column weight-gradient accumulator: acc[mb] (+)= A^T(P) . B, both MN-major, K = the tile's 128 samples
        auto issue_wgrad16 = [&](uint32_t col, uint64_t dB, uint32_t acc_w) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const uint64_t a0 = (pass == 0 ? dPloT : dPhiT) + (uint64_t)((uint32_t)mb * 2u * kSlab >> 4);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) mma_f16(T + col + 16u * mb, a0 + (uint64_t)(kk * 128), dB + (uint64_t)(kk * 128), idW, (pass | kk) ? 1u : acc_w);
                }
        };
        // one GEMM phase: acc (+)= P . B over the 16 ring stages (B MN-major, N = 256, two k-steps per stage); `first_acc` = accumulate flag of
        // the phase's first MMA; the last stage also commits `done`
        auto gemm_phase = [&](uint32_t first_acc, uint64_t* done, unsigned char* gdst) {
#pragma unroll 1
            for (int s = 0; s < 16; ++s, ++q) {
#ifndef GS_WIDE_NOSTORE
                // The tile in P (h1, then dz2) leaves for the workspace in eight 16 KB pieces, one every two stages (P_lo directly follows
                // P_hi): all eight at the phase's start cost the MMAs beside them 3 % more (shared-memory bandwidth; measured).
                if (lane == 0 && (s & 1) == 0) {
                    bulk_s2g(gdst + (uint32_t)(s >> 1) * 16384u, sPhi + (uint32_t)(s >> 1) * 16384u, 16384u);
                    if (s == 14) bulk_commit();
                }
                __syncwarp();
#endif
                const uint32_t slot = q % kRing, round = q / kRing;
                const int sr = (s + rot) & 15;
                const bool lo = sr < 8;
                const uint32_t blk = (uint32_t)sr & 7u;                           // K index 32 blk ..: slab blk / 2 of P, byte 64 (blk & 1) of its rows
                const uint64_t dB = desc(sRing + slot * kStage, 4096u);
                const uint64_t aoff = (uint64_t)(((blk >> 1) * kSlab + (blk & 1u) * 64u) >> 4);
                const uint32_t later = s ? 1u : first_acc;
                mbar_wait(&bars[BAR_FULL + slot], round & 1u);
                fence_after_sync();
                if (s == 15) {                                                    // the tile's copy has left P before `done` lets the next stage overwrite it
                    if (lane == 0) bulk_wait_read();
                    __syncwarp();
                }
                if (elect_one()) {
                    if (!lo) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) mma_f16(T + cAcc, dPlo + aoff + 2u * kk, dB + (uint64_t)(kk * 128), idFullT, kk ? 1u : later);
                    }
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) mma_f16_ts(T + cAcc, T + cAhi + 16u * blk + 8u * kk, dB + (uint64_t)(kk * 128), idFullT, (kk || !lo) ? 1u : later);
                    mma_commit(&bars[BAR_EMPTY + slot]);
                    if (s == 15) mma_commit(done);
                }
                __syncwarp();
            }
        };
        mbar_wait(&bars[RDY_X], 0); fence_after_sync();
        issue_l1(0);
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const uint32_t p = (uint32_t)i & 1u;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)i * G;
            const uint64_t dXg = desc(sX + 32u * (1u + p));
            const uint64_t dXgT = desc(sX + 32u * (1u + p), kSlab);
            const uint32_t acc_w = (i % kFlushTiles) != 0 ? 1u : 0u;
            // ---- forward: z2 = b2 + h1 . W2^T ----
            GS_WT(1, 0);
            mbar_wait(&bars[RDY_H1], p); fence_after_sync();
            GS_WT(1, 1);
            if (elect_one()) mma_f16(T + cAcc, dXg, dWS + 4u, idFull, 0u);
            __syncwarp();
            gemm_phase(1u, &bars[BAR_Z2], tiles + (size_t)tile * kTileBytes);
            GS_WT(1, 3);
            // ---- heads ----
            mbar_wait(&bars[RDY_H2], p); fence_after_sync();
            GS_WT(1, 4);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) mma_f16(T + cH, dPlo + kfeat(kk), dWSh + (uint64_t)(kk * 128), idHeads, kk ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) mma_f16_ts(T + cH, T + cAhi + 8u * kk, dWSh + (uint64_t)(kk * 128), idHeads, 1u);
                mma_commit(&bars[BAR_OUT]);
            }
            __syncwarp();
            GS_WT(1, 5);
            // ---- dh2 = g16 . Wh ; W-c: dWh^T += h2^T g16 ----
            mbar_wait(&bars[RDY_G], p); fence_after_sync();
            GS_WT(1, 6);
            if (elect_one()) {
                mma_f16(T + cAcc, desc(sX), dWS + 6u, idFull, 0u);
                mma_commit(&bars[BAR_DH2]);
                issue_wgrad16(cWh, dG, acc_w);
                mma_commit(&bars[BAR_WC]);
            }
            __syncwarp();
            GS_WT(1, 7);
            // ---- W-d: db2 += dz2^T [ones] ; dgrad: dh1 = dz2 . W2 ----
            mbar_wait(&bars[RDY_DZ2], p); fence_after_sync();
            GS_WT(1, 8);
            if (elect_one()) issue_wgrad16(cB2, dG, acc_w);
            __syncwarp();
            gemm_phase(0u, &bars[BAR_DH1], tiles + (size_t)tile * kTileBytes + 2u * kTile);
            GS_WT(1, 10);
            // ---- W-b: [dW1 | db1] += dz1^T x16 ; next tile's layer 1 ----
            mbar_wait(&bars[RDY_DZ1], p); fence_after_sync();
            GS_WT(1, 11);
            if (elect_one()) {
                issue_wgrad16(cW1, dXgT, acc_w);
                mma_commit(&bars[BAR_WB]);
            }
            __syncwarp();
            GS_WT(1, 12);
            if (i + 1 < n_my) {
                mbar_wait(&bars[RDY_X], p ^ 1u); fence_after_sync();
                issue_l1(i + 1);
            }
            GS_WT(1, 13);
        }
        if (lane == 0) bulk_wait_all();                                           // the stored tiles are complete before the kernel ends
        return;
    }

    // ============================================ compute warps ==================================================================
    const int quad = warp & 3, cg = warp >> 2;                // TMEM lane quadrant, 64-column slab of the row
    const int row = quad * 32 + lane;                         // sample row of the tile == TMEM lane
    const bool loss_thread = cg == 0;
    unsigned char* Xrow = sm + oX + row * 128;
    const int sw = row & 7;
    const uint32_t my_off = (uint32_t)cg * kSlab + (uint32_t)row * 128u;
    const uint32_t T = tmem + ((uint32_t)(quad * 32) << 16);
    float* out = grad_partials + (size_t)blockIdx.x * pstride;
    const ParamOffsets po = param_offsets(D, H, H, A, m.has_value);
    unsigned char* Phi = sm + oPhi;
    unsigned char* Plo = sm + oPlo;

    float pm[PM_N];
#pragma unroll
    for (int q = 0; q < PM_N; ++q) pm[q] = 0.f;
    float gsum[kNH] = {0.f, 0.f, 0.f, 0.f};
    float zs0 = 0.f, zq0 = 0.f, zs1 = 0.f, zq1 = 0.f;
    uint32_t dead0[2] = {0u, 0u}, dead1[2] = {0u, 0u};        // lane l: dead-sample count of neuron 64 cg + 32 half + l

    auto tile_of = [&](int i) -> int64_t { return (int64_t)blockIdx.x + (int64_t)i * G; };
    auto prefetch = [&](int i, uint32_t off, bool ok) {
        const int gx = 2 * (1 + (i & 1));
        const uint32_t d0 = smem_u32(Xrow + (((gx) ^ sw) << 4)), d1 = smem_u32(Xrow + (((gx + 1) ^ sw) << 4));
        if (ok) {
            const uint4* src = records + (size_t)off * 4;
            cp_async16(d0, src); cp_async16(d1, src + 1);
            cp_async16(smem_u32(Xrow + ((6 ^ sw) << 4)), src + 2); cp_async16(smem_u32(Xrow + ((7 ^ sw) << 4)), src + 3);
        } else {
            *reinterpret_cast<uint4*>(Xrow + ((gx ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(Xrow + (((gx + 1) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
    };
    auto offset_of = [&](int i, bool& ok) -> uint32_t {
        const int64_t pos = tile_of(i) * kRows + row;
        ok = i < n_my && pos < b.n;
        return ok ? (offs ? __ldg(offs + pos) : (uint32_t)pos) : 0u;
    };
    // fold the small weight-gradient accumulators (b2, W1, b1, head rows of the 128-feature block `cg`) into the partial vector
    auto flush = [&](bool first) {
        if (cg < 2) {
            const float invB = ncs[4];
            const int j = 128 * cg + row;
            float e[16], w1[16], wh[16];
            tmem_ld16(T + cB2 + 16 * cg, e);
            tmem_ld16(T + cW1 + 16 * cg, w1);
            tmem_ld16(T + cWh + 16 * cg, wh);
            int64_t idx[13];
            idx[0] = po.b2 + j;
#pragma unroll
            for (int d = 0; d < kMaxD; ++d) idx[1 + d] = d < D ? po.w1 + (int64_t)j * D + d : idx[0];
            idx[8] = po.b1 + j;
#pragma unroll
            for (int r = 0; r < 3; ++r) idx[9 + r] = r < A ? po.wp + (int64_t)r * H + j : idx[0];
            idx[12] = m.has_value ? po.wv + j : idx[0];
            float old[13];
#pragma unroll
            for (int q = 0; q < 13; ++q) old[q] = !first ? __ldcg(out + idx[q]) : 0.f;
            tmem_ld_wait();
            float val[13];
            val[0] = e[14];
#pragma unroll
            for (int d = 0; d < kMaxD; ++d) val[1 + d] = w1[d] + w1[7 + d];
            val[8] = w1[14];
#pragma unroll
            for (int r = 0; r < 3; ++r) val[9 + r] = wh[r] + wh[4 + r];
            val[12] = A == 2 ? wh[2] + wh[6] : wh[3] + wh[7];
#pragma unroll
            for (int q = 12; q >= 0; --q) {                   // q = 0 (b2) last: the unused slots alias it
                const bool used = q == 0 || q == 8 || (q >= 1 && q <= 7 && q - 1 < D) || (q >= 9 && q <= 11 && q - 9 < A) || (q == 12 && m.has_value);
                if (used) out[idx[q]] = fmaf(val[q], invB, old[q]);
            }
        }
        fence_before_sync();
    };
    auto warp_ready = [&](int which) {
        tmem_st_wait();
        fence_proxy_async();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[which]);
    };
    // 16 columns of the row as (hi, lo) -> P
    auto store_q = [&](int qt, const uint32_t (&hw)[8], const uint32_t (&lw)[8], bool to_tmem) {
        // the hi half also goes to tensor memory (two fp16 per column): the K-major uses of the tile (forward, dgrad, heads) take A_hi from
        // there, which spares shared memory -- the kernel's bottleneck -- a third of their operand reads
        if (to_tmem) tmem_st8u(T + cAhi + 32 * cg + 8 * qt, hw);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t o = my_off + (uint32_t)(((2 * qt + c) ^ sw) << 4);
            *reinterpret_cast<uint4*>(Phi + o) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
            *reinterpret_cast<uint4*>(Plo + o) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
        }
    };
    // forward stage: the row's 64 pre-activations -> relu -> (hi, lo) -> P, 16 columns at a time (the whole row in registers spilled);
    // returns the relu mask
    auto fwd_stage = [&](bool valid, float& zs, float& zq, uint32_t (&dcnt)[2], uint64_t* release, uint32_t parity) -> uint64_t {   // h1 / h2: both feed K-major MMAs
        uint64_t mask = 0ull;
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
            float z[16];
            tmem_ld16(T + cAcc + 64 * cg + 16 * qt, z);
            tmem_ld_wait();
            if (TRACK) {
                float mn = 1.0f;
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) { zs += z[q]; zq = fmaf(z[q], z[q], zq); mn = fminf(mn, fabsf(z[q])); }
                }
                if (__any_sync(0xffffffffu, mn < 1e-6f)) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const uint32_t hits = __popc(__ballot_sync(0xffffffffu, valid && fabsf(z[q]) < 1e-6f));
                        if (lane == 16 * (qt & 1) + q) dcnt[qt >> 1] += hits;
                    }
                }
            }
            uint32_t bits = 0u;
#pragma unroll
            for (int q = 0; q < 16; ++q) bits |= (z[q] > 0.f ? 1u : 0u) << q;
            mask |= (uint64_t)bits << (16 * qt);
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_pair(fmaxf(z[2 * e], 0.f), fmaxf(z[2 * e + 1], 0.f), hw[e], lw[e]);
            if (qt == 0 && release) mbar_wait(release, parity);
            store_q(qt, hw, lw, true);
        }
        return mask;
    };
    // backward stage: d(loss)/d(activation) of the row * relu' -> (hi, lo) -> P
    auto bwd_stage = [&](uint64_t mask, uint64_t* release, uint32_t parity, bool to_tmem) {
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
            float d[16];
            tmem_ld16(T + cAcc + 64 * cg + 16 * qt, d);
            tmem_ld_wait();
            const uint32_t bits = (uint32_t)(mask >> (16 * qt));
#pragma unroll
            for (int q = 0; q < 16; ++q) d[q] = (bits >> q) & 1u ? d[q] : 0.f;
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_pair(d[2 * e], d[2 * e + 1], hw[e], lw[e]);
            if (qt == 0 && release) mbar_wait(release, parity);
            store_q(qt, hw, lw, to_tmem);
        }
    };

    uint32_t next_off = 0;
    bool next_ok = false;
    if (loss_thread) {
        bool ok;
        const uint32_t off = offset_of(0, ok);
        prefetch(0, off, ok);
        next_off = offset_of(1, next_ok);
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[RDY_X]);
    }
#pragma unroll 1
    for (int i = 0; i < n_my; ++i) {
        const uint32_t p = (uint32_t)i & 1u;
        const bool valid = tile_of(i) * kRows + row < b.n;
        // ---- A: h1 ----------------------------------------------------------------------------------------------------------
        GS_WT(0, 0);
        mbar_wait(&bars[BAR_Z1], p);
        fence_after_sync();
        GS_WT(0, 1);
        if (i > 0 && (i % kFlushTiles) == 0) {               // every weight-gradient MMA of the previous tiles has completed
            mbar_wait(&bars[BAR_WB], p ^ 1u);
            fence_after_sync();
            flush(i == kFlushTiles);
        }
        const uint64_t mask1 = fwd_stage(valid, zs0, zq0, dead0, i > 0 ? &bars[BAR_WB] : nullptr, p ^ 1u);   // W-b(i-1) reads dz1 from P
        warp_ready(RDY_H1);
        GS_WT(0, 2);
        // ---- B: h2 ----------------------------------------------------------------------------------------------------------
        mbar_wait(&bars[BAR_Z2], p);                          // every forward MMA is done and the h1 tile has been copied out of P
        fence_after_sync();
        GS_WT(0, 3);
        const uint64_t mask2 = fwd_stage(valid, zs1, zq1, dead1, nullptr, 0u);
        warp_ready(RDY_H2);
        GS_WT(0, 4);
        // ---- C: loss (one thread per row) ----------------------------------------------------------------------------------------
        if (loss_thread) {
            mbar_wait(&bars[BAR_OUT], p);
            fence_after_sync();
            GS_WT(0, 5);
            float c[16];
            tmem_ld16(T + cH, c);
            const uint4 sc = *reinterpret_cast<const uint4*>(Xrow + ((6 ^ sw) << 4));
            const float ret_s = *reinterpret_cast<const float*>(Xrow + ((7 ^ sw) << 4));
            tmem_ld_wait();
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
            uint32_t h01, l01, h23, l23;
            split_pair(g[0], g[1], h01, l01);
            split_pair(g[2], g[3], h23, l23);
            *reinterpret_cast<uint4*>(Xrow + ((0 ^ sw) << 4)) = make_uint4(h01, h23, l01, l23);
            *reinterpret_cast<uint4*>(Xrow + ((1 ^ sw) << 4)) = make_uint4(h01, h23, 0u, kOnes2);
            warp_ready(RDY_G);
            GS_WT(0, 6);
            if (i + 1 < n_my) prefetch(i + 1, next_off, next_ok);   // off the critical path; the scalars of this tile were consumed above
            next_off = offset_of(i + 2, next_ok);
        }
        // ---- D: dz2 (over h2, once W-c has read it) ----------------------------------------------------------------------------------
        mbar_wait(&bars[BAR_DH2], p);
        fence_after_sync();
        GS_WT(0, 7);
        bwd_stage(mask2, &bars[BAR_WC], p, true);       // dz2 feeds dgrad (K-major A)
        warp_ready(RDY_DZ2);
        GS_WT(0, 8);
        // ---- E: dz1 (over dz2: dgrad, W-d and the tile store are done once dh1 is complete) ---------------------------------------------
        if (loss_thread && i + 1 < n_my) {                    // the next tile's record has landed: its layer 1 may start after W-b
            cp_async_wait_all();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[RDY_X]);
        }
        mbar_wait(&bars[BAR_DH1], p);
        fence_after_sync();
        GS_WT(0, 9);
        bwd_stage(mask1, nullptr, 0u, false);           // dz1 is only read MN-major (W-b)
        warp_ready(RDY_DZ1);
        GS_WT(0, 10);
    }
    mbar_wait(&bars[BAR_WB], (uint32_t)(n_my - 1) & 1u);
    fence_after_sync();
    flush(n_my <= kFlushTiles);
    // ---- CTA reductions: metric partials, activation statistics, head-bias gradients ---------------------------------------------
    if (loss_thread) {
#pragma unroll
        for (int q = 0; q < PM_N; ++q) {
            const float v = warp_sum(pm[q]);
            if (lane == 0 && v != 0.f) atomicAdd(red + q, (double)v);
        }
#pragma unroll
        for (int r = 0; r < kNH; ++r) {
            const float v = warp_sum(gsum[r]);
            if (lane == 0) fr[quad * 4 + r] = v;
        }
    }
    if (TRACK) {
        const double a0 = warp_sum((double)zs0), a1 = warp_sum((double)zq0), a2 = warp_sum((double)zs1), a3 = warp_sum((double)zq1);
        if (lane == 0) { atomicAdd(red + PM_Z0, a0); atomicAdd(red + PM_Z0SQ, a1); atomicAdd(red + PM_Z1, a2); atomicAdd(red + PM_Z1SQ, a3); }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (dead0[half]) atomicAdd(dead + 64 * cg + 32 * half + lane, dead0[half]);
            if (dead1[half]) atomicAdd(dead + H + 64 * cg + 32 * half + lane, dead1[half]);
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory");
    if (tid < PM_N) metric_partials[(size_t)blockIdx.x * PM_N + tid] = red[tid];
    if (warp == 0 && lane < kNH) {                            // fixed order: deterministic
        const float sgm = (fr[0 * 4 + lane] + fr[1 * 4 + lane]) + (fr[2 * 4 + lane] + fr[3 * 4 + lane]);
        if (lane < A) out[po.bp + lane] = sgm * ncs[4];
        else if (lane == A && m.has_value) out[po.bv] = sgm * ncs[4];
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// dW2 = (1/B) dz2^T . h1 over the tiles the fused kernel stored: M = 2 x 128 (j), N = 256 (k), K = samples; one CTA per SM, CTA c
// takes tiles c, c + G, .. and writes the w2 block of partial vector c (the vector whose other entries update_wide_kernel's CTA c wrote).
__global__ void __launch_bounds__(192, 1)
wgrad_wide_kernel(const unsigned char* __restrict__ tiles, int64_t n, float* __restrict__ grad_partials, int64_t pstride, int64_t w2_off) {
    using namespace wfu;
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kGRing * kGStage);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kGRing * kGStage + 8 * kGBars);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int k = 0; k < kGBars; ++k) mbar_init(&bars[k], k == G_FOLDED ? 4u : 1u);
        fence_mbar_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(*tmem_slot);
    const uint32_t warp_u = uniform((uint32_t)warp);
    const int64_t n_tiles = (n + kRows - 1) / kRows;
    const int G = (int)gridDim.x;
    const int n_my = (int)((n_tiles - blockIdx.x + G - 1) / G);
    const uint32_t sRing = smem_u32(sm);

    if (warp_u == 4u) {                                       // loader
        if (lane == 0) {
            uint32_t q = 0;
#pragma unroll 1
            for (int i = 0; i < n_my; ++i) {
                const unsigned char* src = tiles + (size_t)((int64_t)blockIdx.x + (int64_t)i * G) * kTileBytes;
#pragma unroll 1
                for (int s = 0; s < 4; ++s, ++q) {
                    const uint32_t slot = q % kGRing, round = q / kGRing;
                    mbar_wait(&bars[G_EMPTY + slot], (round & 1u) ^ 1u);
                    mbar_expect_tx(&bars[G_FULL + slot], kGStage);
#pragma unroll 1
                    for (uint32_t pi = 0; pi < 16u; ++pi)     // piece pi = (tensor-precision pi / 4, slab pi % 4): rows 32 s .. of that slab
                        bulk_g2s(sRing + slot * kGStage + pi * kGPiece, src + (size_t)pi * kSlab + (size_t)s * kGPiece, kGPiece, &bars[G_FULL + slot]);
                }
            }
        }
    } else if (warp_u == 5u) {                                // MMA issue
        if (lane == 0) {
            uint32_t q = 0;
#pragma unroll 1
            for (int i = 0; i < n_my; ++i) {
                if (i > 0 && (i % kFlushTiles) == 0) {        // the epilogue warps fold the accumulators, then they restart
                    mma_commit(&bars[G_ACC]);
                    mbar_wait(&bars[G_FOLDED], (uint32_t)(i / kFlushTiles - 1) & 1u);
                    fence_after_sync();
                }
                const bool restart = (i % kFlushTiles) == 0;
#pragma unroll 1
                for (int s = 0; s < 4; ++s, ++q) {
                    const uint32_t slot = q % kGRing, round = q / kGRing;
                    const uint32_t sS = sRing + slot * kGStage;
                    mbar_wait(&bars[G_FULL + slot], round & 1u);
                    fence_after_sync();
#pragma unroll 1
                    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {                    // dz2_lo h1_hi, dz2_hi h1_lo, dz2_hi h1_hi
                            const uint32_t a0 = sS + (pass == 0 ? 12u : 8u) * kGPiece + (uint32_t)mb * 2u * kGPiece;
                            const uint32_t b0 = sS + (pass == 1 ? 4u : 0u) * kGPiece;
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                mma_f16(tmem + 256u * mb, desc(a0 + kk * 2048u, kGPiece), desc(b0 + kk * 2048u, kGPiece), idesc_f16(128, 256, 1, 1),
                                        (restart && s == 0 && pass == 0 && kk == 0) ? 0u : 1u);
                        }
                    mma_commit(&bars[G_EMPTY + slot]);
                }
            }
            mma_commit(&bars[G_ACC]);
        }
    } else {                                                  // warps 0..3: fold the accumulators into the partial vector
        const int row = warp * 32 + lane;
        const uint32_t T = tmem + ((uint32_t)(warp * 32) << 16);
        const float invB = 1.0f / (float)n;
        float* out = grad_partials + (size_t)blockIdx.x * pstride + w2_off;
        const int n_folds = (n_my + kFlushTiles - 1) / kFlushTiles;
#pragma unroll 1
        for (int f = 0; f < n_folds; ++f) {
            mbar_wait(&bars[G_ACC], (uint32_t)f & 1u);
            fence_after_sync();
#pragma unroll 1
            for (int mb = 0; mb < 2; ++mb) {
                float* d = out + (int64_t)(128 * mb + row) * H;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    float v[32];
                    tmem_ld32(T + 256u * mb + 32u * c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) d[32 * c + e] = f == 0 ? v[e] * invB : fmaf(v[e], invB, d[32 * c + e]);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0 && f + 1 < n_folds) mbar_arrive(&bars[G_FOLDED]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_stage_w2(const MlpDev& md, unsigned char* staged, cudaStream_t st) {
    stage_w2_kernel<<<32, 256, 0, st>>>(md, staged);
    GS_LAUNCH_CHECK();
    return 0;
}

// ---- host launcher (called from update_kernels.cu) -------------------------------------------------------------------------------
int64_t wide_scratch_bytes(int64_t max_batch) {
    const int64_t tiles = (max_batch + 127) / 128;
    return 4ll * wfu::kW2Prec + tiles * (int64_t)wfu::kTileBytes;
}

template <int ALGO>
int launch_update_wide(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                       const uint32_t* offs, const void* records, void* scratch, float* grad_partials, int64_t pstride, double* metric_partials,
                       uint32_t* dead, int grid, cudaStream_t st) {
    using namespace wfu;
    unsigned char* w2s = reinterpret_cast<unsigned char*>(scratch);
    unsigned char* tiles = w2s + 4 * kW2Prec;
    if (launch_stage_w2(md, w2s, st)) return -1;
    using KernelFn = void (*)(MlpDev, BatchDev, HpDev, const double*, const double*, const uint32_t*, const uint4*, const unsigned char*, unsigned char*,
                              float*, int64_t, double*, uint32_t*);
    KernelFn kern = track ? update_wide_kernel<ALGO, true> : update_wide_kernel<ALGO, false>;
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    kern<<<grid, kWideThreads, kSmemBytes, st>>>(md, b, hp, adv_mom, ret_mom, offs, reinterpret_cast<const uint4*>(records), w2s, tiles, grad_partials,
                                             pstride, metric_partials, dead);
    GS_LAUNCH_CHECK();
    GS_CUDA(cudaFuncSetAttribute(wgrad_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGSmem));
    const ParamOffsets po = param_offsets(md.D, H, H, md.A, md.has_value);
    wgrad_wide_kernel<<<grid, 192, kGSmem, st>>>(tiles, b.n, grad_partials, pstride, po.w2);
    GS_LAUNCH_CHECK();
    return 0;
}
template int launch_update_wide<ALGO_PPO>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, const uint32_t*, const void*, void*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);
template int launch_update_wide<ALGO_REINFORCE>(const MlpDev&, const BatchDev&, const HpDev&, bool, const double*, const double*, const uint32_t*, const void*, void*, float*, int64_t, double*, uint32_t*, int, cudaStream_t);

}  // namespace gs

#ifdef GS_WIDE_TRACE
extern "C" int gs_debug_wide_trace(long long* host_out /* [2][24] */) {
    return (int)cudaMemcpyFromSymbol(host_out, gs::g_wide_trace, sizeof(long long) * 48);
}
#endif
