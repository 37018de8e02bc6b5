// collect_wide.cu — the fused collect kernel and policy_act for the 256 x 256 MLP (`mlp_medium`) with the forward pass on the tensor cores
// (fp16x3).  Same contract as collect_f16.cu (utils/policy_ops.py:14-41, utils/rollout_collector.py:474-542, utils/rollout_buffer.py:82-102)
// and the same per-step chain  x16(obs) -> L1 -> act, split -> h1 tile -> bias + W2 -> act, split -> h2 tile -> heads -> action / env step;
// what changes at H = 256 is what update_wide.cu describes: W2^T as operand tiles (hi + lo = 256 KB) does not fit shared memory, so it is
// staged once per call in global memory (wide.cuh::launch_stage_w2) and streamed through a 3 x 16 KB cp.async.bulk ring by a loader warp,
// 16 stages per step; one activation buffer holds h1, then h2.  A CTA = 128 environments (TMEM lanes) carried for all T steps with their fp64
// physics state in registers (16 compute warps: thread = env row x 64-column slab; the warps of slab 0 are the env threads), one CTA per SM.
// MODE_ACT is the same kernel over an observation array (gs_policy_act / gs_policy_values), so fused and unfused sequences stay bit-identical.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "wide.cuh"
#include "rollout_shared.cuh"

namespace gs {

using namespace tc;

namespace cwu {

using namespace wfu;   // tile / ring geometry, bulk-copy helpers (and hfu underneath)

enum { CBAR_Z1 = 0, CBAR_Z2, CBAR_OUT, CRDY_X, CRDY_H1, CRDY_H2, CBAR_FULL, CBAR_EMPTY = CBAR_FULL + kRing, kCBars = CBAR_EMPTY + kRing };
enum { MODE_COLLECT = 0, MODE_ACT = 1 };
constexpr uint32_t oBars = oRing + kRing * kStage, oTmem = oBars + 8 * kCBars, oBH = oTmem + 16, kSmemBytes = oBH + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget");
constexpr uint32_t cAcc = 0, cH = 256, kCols = 512;

struct ActDev {   // MODE_ACT arguments (policy_act over an observation array)
    const float* obs;
    int64_t n, row_offset;
    const float* uniforms;
    int32_t* actions;
    float *logp, *value, *logits_out;
};

}  // namespace cwu

template <int KIND, int MODE>
__global__ void __launch_bounds__(wfu::kWideThreads, 1)
collect_wide_kernel(EnvDev h, MlpDev m, RolloutDev buf, float* __restrict__ cur_obs, cwu::ActDev act, const unsigned char* __restrict__ w2t /* W2^T hi, lo */,
                    uint64_t rng_seed, uint64_t step0, int deterministic) {
    using namespace cwu;
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + cwu::oBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + cwu::oTmem);
    float* bhs = reinterpret_cast<float*>(sm + cwu::oBH);
    constexpr int D = MODE == MODE_COLLECT ? EnvDims<KIND>::D : kMaxD;     // MODE_ACT: runtime m.D <= 7

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = m.A;
    if (warp == 0) tmem_alloc(tmem_slot, cwu::kCols);
    if (tid == 32) {
        for (int k = 0; k < kCBars; ++k)
            mbar_init(&bars[k], (k < CRDY_X || k >= CBAR_FULL) ? 1u : (k == CRDY_X ? 4u : (uint32_t)kCW));
        fence_mbar_init();
    }
    {
        uint4* x4 = reinterpret_cast<uint4*>(sm + oX);
        for (int i = tid; i < (int)(kSlab / 16); i += kWideThreads) x4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    stage_ws<H>(m, sm + oWS, tid);
    if (tid < 4) bhs[tid] = tid < A ? __ldg(m.bp + tid) : ((tid == A && m.has_value) ? __ldg(m.bv) : 0.f);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(*tmem_slot);
    const uint32_t warp_u = uniform((uint32_t)warp);

    // iterations of the forward pass: T steps + V(last_obs) (collect), or this CTA's 128-row groups (act)
    int64_t n_iter;
    if (MODE == MODE_COLLECT) n_iter = (int64_t)buf.T + 1;
    else {
        const int64_t n_groups = (act.n + kRows - 1) / kRows;
        n_iter = (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x;
    }
    const uint32_t sBase = smem_u32(sm);
    const uint32_t sPhi = sBase + oPhi, sPlo = sBase + oPlo, sX = sBase + oX, sWS = sBase + oWS, sRing = sBase + oRing;
    const int rot = (int)(blockIdx.x & 15u);

    // ============================================ loader warp: W2^T blocks -> ring ==============================================
    if (warp_u == (uint32_t)kCW + 1u) {
        if (lane == 0) {
            uint32_t q = 0;
#pragma unroll 1
            for (int64_t it = 0; it < n_iter; ++it) {
#pragma unroll 1
                for (int s = 0; s < 16; ++s, ++q) {
                    const uint32_t slot = q % kRing, round = q / kRing;
                    mbar_wait(&bars[CBAR_EMPTY + slot], (round & 1u) ^ 1u);
                    mbar_expect_tx(&bars[CBAR_FULL + slot], kStage);
                    const int s16 = (s + rot) & 15;
                    const uint32_t prec = s16 < 8 ? 1u : 0u, blk = (uint32_t)s16 & 7u;
                    const unsigned char* src = w2t + prec * kW2Prec + blk * 4096u;
                    const uint32_t dst = sRing + slot * kStage;
#pragma unroll
                    for (uint32_t sl = 0; sl < 4u; ++sl) bulk_g2s(dst + sl * 4096u, src + sl * (H * 128u), 4096u, &bars[CBAR_FULL + slot]);
                }
            }
        }
        return;
    }

    // ============================================ MMA-issuing warp ==============================================================
    if (warp_u == (uint32_t)kCW) {
        const uint32_t T = tmem;
        uint32_t q = 0;
        const uint64_t dPhi = desc(sPhi), dPlo = desc(sPlo), dX = desc(sX);
        const uint64_t dWS = desc(sWS), dWSh = desc(sWS + 96u, H * 128u);
        constexpr uint32_t idFull = idesc_f16(128, H, 0, 0), idFullT = idesc_f16(128, H, 0, 1), idHeads = idesc_f16(128, 16, 0, 1);
        auto kfeat = [](int kk) -> uint64_t { return (uint64_t)(((uint32_t)(kk >> 2) * kSlab + (uint32_t)(kk & 3) * 32u) >> 4); };
#pragma unroll 1
        for (int64_t it = 0; it < n_iter; ++it) {
            const uint32_t p = (uint32_t)it & 1u;
            mbar_wait(&bars[CRDY_X], p); fence_after_sync();
            if (elect_one()) {
                mma_f16(T + cwu::cAcc, dX, dWS + 2u, idFull, 0u);
                mma_f16(T + cwu::cAcc, dX, dWS, idFull, 1u);
                mma_commit(&bars[CBAR_Z1]);
            }
            __syncwarp();
            mbar_wait(&bars[CRDY_H1], p); fence_after_sync();
            if (elect_one()) mma_f16(T + cwu::cAcc, dX, dWS + 4u, idFull, 0u);
            __syncwarp();
#pragma unroll 1
            for (int s = 0; s < 16; ++s, ++q) {
                const uint32_t slot = q % kRing, round = q / kRing;
                const int sr = (s + rot) & 15;
                const bool lo = sr < 8;
                const uint32_t blk = (uint32_t)sr & 7u;
                const uint64_t dB = desc(sRing + slot * kStage, 4096u);
                const uint64_t aoff = (uint64_t)(((blk >> 1) * kSlab + (blk & 1u) * 64u) >> 4);
                mbar_wait(&bars[CBAR_FULL + slot], round & 1u);
                fence_after_sync();
                if (elect_one()) {
                    if (!lo) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) mma_f16(T + cwu::cAcc, dPlo + aoff + 2u * kk, dB + (uint64_t)(kk * 128), idFullT, 1u);
                    }
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) mma_f16(T + cwu::cAcc, dPhi + aoff + 2u * kk, dB + (uint64_t)(kk * 128), idFullT, 1u);
                    mma_commit(&bars[CBAR_EMPTY + slot]);
                    if (s == 15) mma_commit(&bars[CBAR_Z2]);
                }
                __syncwarp();
            }
            mbar_wait(&bars[CRDY_H2], p); fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const uint64_t a0 = pass == 0 ? dPlo : dPhi;
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) mma_f16(T + cwu::cH, a0 + kfeat(kk), dWSh + (uint64_t)(kk * 128), idHeads, (pass | kk) ? 1u : 0u);
                }
                mma_commit(&bars[CBAR_OUT]);
            }
            __syncwarp();
        }
        return;
    }

    // ============================================ compute warps ==================================================================
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const bool env_thread = cg == 0;                          // the warps of slab 0 own the environments / rows
    const int sw = row & 7;
    const uint32_t my_off = (uint32_t)cg * kSlab + (uint32_t)row * 128u;
    const uint32_t T = tmem + ((uint32_t)(quad * 32) << 16);
    unsigned char* Xrow = sm + oX + row * 128;
    unsigned char* Phi = sm + oPhi;
    unsigned char* Plo = sm + oPlo;

    // accumulator -> activation -> (hi, lo) -> P, 16 columns at a time
    auto stage = [&]() {
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
            float z[16];
            tmem_ld16(T + cwu::cAcc + 64 * cg + 16 * qt, z);
            tmem_ld_wait();
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_pair(act_fwd(z[2 * e], m.act), act_fwd(z[2 * e + 1], m.act), hw[e], lw[e]);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t o = my_off + (uint32_t)(((2 * qt + c) ^ sw) << 4);
                *reinterpret_cast<uint4*>(Phi + o) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
                *reinterpret_cast<uint4*>(Plo + o) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
            }
        }
    };
    auto warp_ready = [&](uint64_t* bar) {
        fence_proxy_async();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    };
    // x16 = [x_hi | x_lo | 1 1] of one observation -> 16-column group 0 of the row (zeros for rows outside the batch)
    auto put_x16 = [&](const float* o, int d_used, bool ok) {
        uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (ok) {
#pragma unroll
            for (int d = 0; d < kMaxD; ++d)
                if (d < D && d < d_used) {
                    const uint32_t hb = f16_bits(o[d]);
                    float xh, dummy;
                    unpack_pair(hb, xh, dummy);
                    const uint32_t lb = f16_bits(o[d] - xh);
                    w[d >> 1] |= hb << ((d & 1) * 16);
                    w[(7 + d) >> 1] |= lb << (((7 + d) & 1) * 16);
                }
            w[7] = kOnes2;
        }
        *reinterpret_cast<uint4*>(Xrow + ((0 ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(Xrow + ((1 ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
    };

    // ---- env-thread state ---------------------------------------------------------------------------------------------------
    const int64_t n_env = MODE == MODE_COLLECT ? (int64_t)blockIdx.x * kRows + row : 0;
    const bool owner = MODE == MODE_COLLECT && env_thread && n_env < h.n;
    EnvRegs e;
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    if (owner) {
        env_load<KIND>(h, n_env, e);
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = cur_obs[n_env * D + d];
    }
    const uint64_t gid = (uint64_t)(h.params.gid0 + n_env);
    int64_t arow = 0;
    bool aok = false;

    auto produce = [&](int64_t it) {
        if (MODE == MODE_COLLECT) put_x16(o, D, owner);
        else {
            arow = ((int64_t)blockIdx.x + it * gridDim.x) * kRows + row;
            aok = arow < act.n;
#pragma unroll
            for (int d = 0; d < D; ++d) o[d] = (aok && d < m.D) ? __ldg(act.obs + arow * m.D + d) : 0.f;
            put_x16(o, m.D, aok);
        }
        warp_ready(&bars[CRDY_X]);                             // (its tcgen05 fence also orders the read of the previous head outputs)
    };
    auto consume = [&](int64_t it) {
        mbar_wait(&bars[CBAR_OUT], (uint32_t)it & 1u);
        fence_after_sync();
        float c[16];
        tmem_ld16(T + cwu::cH, c);
        tmem_ld_wait();
        float out[kNH];
#pragma unroll
        for (int r = 0; r < kNH; ++r) out[r] = bhs[r] + (c[r] + c[8 + r]);
        if (MODE == MODE_COLLECT) {
            if (!owner) return;
            if (it < buf.T) {
                const int t = (int)it;
                const float u = deterministic ? 0.f : action_uniform(rng_seed, gid, step0 + (uint64_t)t);
                int a; float lp, v;
                act_from_heads(out, A, m.has_value, deterministic != 0, u, a, lp, v);
                const int64_t off = (int64_t)t * buf.N + n_env;
                store_obs_row<D>(buf.obs, off, o);
                buf.actions[off] = a;
                buf.logprobs[off] = lp;
                buf.values[off] = v;
                double r, ep_r;
                bool term, trunc;
                int ep_l;
                env_vec_step<KIND>(e, h.params, n_env, a, o, r, term, trunc, ep_r, ep_l);
                buf.rewards[off] = (float)r;
                buf.dones[off] = (term || trunc) ? 1 : 0;
                buf.timeouts[off] = trunc ? 1 : 0;
                if (buf.next_obs) store_obs_row<D>(buf.next_obs, off, o);
                if (buf.ep_return) buf.ep_return[off] = ep_r;
                if (buf.ep_length) buf.ep_length[off] = ep_l;
            } else {                                   // V(last_obs) for the GAE bootstrap (rollout_collector.py:373) and the state hand-back
                if (buf.last_values) buf.last_values[n_env] = m.has_value ? (A == 2 ? out[2] : out[3]) : 0.f;
                if (buf.last_obs) store_obs_row<D>(buf.last_obs, n_env, o);
                store_obs_row<D>(cur_obs, n_env, o);
                env_store<KIND>(h, n_env, e);
            }
        } else if (aok) {
            if (act.actions) {
                const float u = deterministic ? 0.f : (act.uniforms ? __ldg(act.uniforms + arow) : action_uniform(rng_seed, (uint64_t)(act.row_offset + arow), step0));
                int a; float lp, v;
                act_from_heads(out, A, m.has_value, deterministic != 0, u, a, lp, v);
                act.actions[arow] = a;
                if (act.logp) act.logp[arow] = lp;
                if (act.value) act.value[arow] = v;
            } else if (act.value) {
                act.value[arow] = m.has_value ? (A == 2 ? out[2] : out[3]) : 0.f;
            }
            if (act.logits_out)
                for (int k = 0; k < A; ++k) act.logits_out[arow * A + k] = k == 0 ? out[0] : (k == 1 ? out[1] : out[2]);
        }
    };

#pragma unroll 1
    for (int64_t it = 0; it < n_iter; ++it) {
        const uint32_t p = (uint32_t)it & 1u;
        if (env_thread) {
            if (it > 0) consume(it - 1);
            produce(it);
        }
        mbar_wait(&bars[CBAR_Z1], p);                          // (transitively: the heads MMAs of the previous step are done with P)
        fence_after_sync();
        stage();
        warp_ready(&bars[CRDY_H1]);
        mbar_wait(&bars[CBAR_Z2], p);
        fence_after_sync();
        stage();
        warp_ready(&bars[CRDY_H2]);
    }
    if (env_thread && n_iter > 0) consume(n_iter - 1);
    fence_before_sync();
    asm volatile("bar.sync 1, %0;" ::"n"(wfu::kCompute) : "memory");
    if (warp == 0) tmem_dealloc(tmem, cwu::kCols);
}

// ---- host launchers (called from collect_f16.cu's dispatch) --------------------------------------------------------------------
template <int KIND, int MODE>
static int launch_wide(const EnvDev& h, const MlpDev& md, const RolloutDev& buf, float* cur_obs, const cwu::ActDev& act, unsigned grid, uint64_t seed,
                       uint64_t step0, int deterministic, cudaStream_t st) {
    // W2 / W2^T in the operand layout, staged once per call (the weights are fixed during a rollout) into a 1 MB buffer kept per
    // (device, stream): calls on one stream are ordered, calls on different streams (training and background evaluation) get their own
    unsigned char* staged = nullptr;
    {
        static std::mutex mu;
        static std::map<std::pair<int, cudaStream_t>, unsigned char*> cache;
        int device = 0;
        GS_CUDA(cudaGetDevice(&device));
        std::lock_guard<std::mutex> lock(mu);
        unsigned char*& slot = cache[std::make_pair(device, st)];
        if (!slot) GS_CUDA(cudaMalloc((void**)&slot, 4 * (size_t)wfu::kW2Prec));
        staged = slot;
    }
    if (launch_stage_w2(md, staged, st)) return -1;
    auto kern = collect_wide_kernel<KIND, MODE>;
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cwu::kSmemBytes));
    kern<<<grid, wfu::kWideThreads, cwu::kSmemBytes, st>>>(h, md, buf, cur_obs, act, staged + 2 * (size_t)wfu::kW2Prec, seed, step0, deterministic);
    GS_LAUNCH_CHECK();
    return 0;
}

int launch_collect_wide(gs_env* env, const MlpDev& md, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0, int deterministic,
                        cudaStream_t st) {
    const unsigned grid = (unsigned)((env->n + hfu::kRows - 1) / hfu::kRows);
    const EnvDev h = to_dev(env);
    if (env->kind == GS_ENV_CARTPOLE_V1) return launch_wide<GS_ENV_CARTPOLE_V1, cwu::MODE_COLLECT>(h, md, buf, cur_obs, cwu::ActDev{}, grid, seed, step0, deterministic, st);
    if (env->kind == GS_ENV_ACROBOT_V1) return launch_wide<GS_ENV_ACROBOT_V1, cwu::MODE_COLLECT>(h, md, buf, cur_obs, cwu::ActDev{}, grid, seed, step0, deterministic, st);
    return launch_wide<GS_ENV_MOUNTAINCAR_V0, cwu::MODE_COLLECT>(h, md, buf, cur_obs, cwu::ActDev{}, grid, seed, step0, deterministic, st);
}

int launch_policy_act_wide(const MlpDev& md, const float* obs, int64_t n, uint64_t seed, uint64_t offset, int64_t row_offset, int deterministic,
                           const float* uniforms, int32_t* actions, float* logp, float* value, float* logits, cudaStream_t st) {
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    cwu::ActDev act;
    act.obs = obs; act.n = n; act.row_offset = row_offset; act.uniforms = uniforms; act.actions = actions; act.logp = logp; act.value = value;
    act.logits_out = logits;
    const int64_t n_groups = (n + hfu::kRows - 1) / hfu::kRows;
    const int64_t cap = sm_count(device);
    const unsigned grid = (unsigned)(n_groups < cap ? n_groups : cap);
    return launch_wide<GS_ENV_CARTPOLE_V1, cwu::MODE_ACT>(EnvDev{}, md, RolloutDev{}, nullptr, act, grid, seed, offset, deterministic, st);
}

}  // namespace gs
