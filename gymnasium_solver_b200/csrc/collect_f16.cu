// collect_f16.cu — the fused collect kernel and policy_act with the MLP forward on the tensor cores (fp16x3, see f16x3.cuh), for the
// H x H networks update_f16.cu serves (H = 64, 128; observations of up to 7 features); H = 256 is dispatched to collect_wide.cu.
//
// Replaces, like rollout_kernels.cu: utils/policy_ops.py:14-41 (policy_act, policy_predict_values) and the per-step body of
// RolloutCollector._collect (utils/rollout_collector.py:474-542) + RolloutBuffer.add (utils/rollout_buffer.py:82-102).
//
// A SET is 128 environments (TMEM lanes) carried for all T steps; a CTA runs two sets for the 64-wide network (one for 128) with the same
// warps, so one set's MMA groups and hand-overs run under the other set's SIMT stages, and 65,536 environments are ONE wave of 256 CTAs at
// two CTAs per SM.  One "env thread" per environment (the warps of 32-column group s own the environments of set s) keeps the fp64 physics state in registers,
// draws the action and writes the step's obs / action / logp / value / reward / done / timeout straight into the time-major buffer;
// all compute warps (thread = env row x 32 hidden units) turn the accumulators into the next operand tile; one warp issues the MMAs:
//   x16(obs) -> L1 (2 MMAs) -> relu, split -> h1 tile -> bias + W2 (1 + 3 H/16 MMAs) -> relu, split -> h2 tile -> heads (2 H/16 MMAs)
// per step, handed over through mbarriers exactly as in the update kernel (no CTA-wide barrier in the loop).  The same kernel without the
// environments (MODE_ACT) is gs_policy_act / gs_policy_values over an observation array, so a fused collect and the unfused sequence
// policy_act + env_step stay bit-identical.
#include <cstdlib>
#include <cstring>

#include "f16x3.cuh"
#include "rollout_shared.cuh"

namespace gs {

using namespace tc;

namespace cfu {

using namespace hfu;

enum { BAR_Z1 = 0, BAR_Z2, BAR_OUT, RDY_X, RDY_H1, RDY_H2, kBars };
enum { MODE_COLLECT = 0, MODE_ACT = 1 };

template <int H>
struct Cfg {
    static constexpr int kSets = H == 64 ? 2 : 1;
    static constexpr int kCG = H / 32, kCW = 4 * kCG, kCompute = kCW * 32, kThreads = kCompute + 32 * kSets;
    static constexpr int kSlabs = H / 64;
    static constexpr uint32_t kTile = kSlabs * kSlab;
    static constexpr uint32_t kW2 = H * H * 2, kW2Slab = H * 128;
    // per set: P_hi, P_lo (h1, then h2); shared: X (x16 of set s in 16-column group s), W2 hi / lo, WS
    static constexpr uint32_t kSetBytes = 2 * kTile, oX = kSets * kSetBytes, oW2hi = oX + kSlab, oW2lo = oW2hi + kW2, oWS = oW2lo + kW2, oBars = oWS + H * 128;
    static constexpr uint32_t oTmem = oBars + 8 * kBars * kSets, oBH = oTmem + 16, kSmemBytes = oBH + 16;
    static constexpr uint32_t cAcc = 0, cH = H, kSetCols = H == 64 ? 128 : 256, kCols = kSets * kSetCols;     // z1 / z2 share one accumulator (strictly serial)
    static constexpr int kMinCtas = H == 64 ? 2 : 1;
    static_assert(kSmemBytes * kMinCtas <= 232448 - 1024 * kMinCtas, "shared memory budget");
    static_assert(kCols * kMinCtas <= 512, "TMEM budget");
};

struct ActDev {   // MODE_ACT arguments (policy_act over an observation array)
    const float* obs;
    int64_t n, row_offset;
    const float* uniforms;
    int32_t* actions;
    float *logp, *value, *logits_out;
};

}  // namespace cfu

template <int H, int KIND, int MODE>
__global__ void __launch_bounds__(cfu::Cfg<H>::kThreads, cfu::Cfg<H>::kMinCtas)
collect_f16_kernel(EnvDev h, MlpDev m, RolloutDev buf, float* __restrict__ cur_obs, cfu::ActDev act, uint64_t rng_seed, uint64_t step0, int deterministic) {
    using C = cfu::Cfg<H>;
    using namespace cfu;
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(sm + C::oBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::oTmem);
    float* bhs = reinterpret_cast<float*>(sm + C::oBH);
    constexpr int D = MODE == MODE_COLLECT ? EnvDims<KIND>::D : kMaxD;     // MODE_ACT: runtime m.D <= 7

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = m.A;
    if (warp == 0) tmem_alloc(tmem_slot, C::kCols);
    if (tid == 32) {
        for (int s = 0; s < C::kSets; ++s)
            for (int k = 0; k < kBars; ++k) mbar_init(&bars_all[s * kBars + k], k < RDY_X ? 1u : (k == RDY_X ? 4u : (uint32_t)C::kCW));
        fence_mbar_init();
    }
    {
        uint4* x4 = reinterpret_cast<uint4*>(sm + C::oX);
        for (int i = tid; i < (int)(kSlab / 16); i += C::kThreads) x4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    stage_weights<H>(m, sm + C::oW2hi, sm + C::oW2lo, sm + C::oWS, tid, C::kThreads);
    if (tid < 4) bhs[tid] = tid < A ? __ldg(m.bp + tid) : ((tid == A && m.has_value) ? __ldg(m.bv) : 0.f);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = uniform(*tmem_slot);
    const uint32_t warp_u = uniform((uint32_t)warp);

    // iterations of the forward pass: T steps + V(last_obs) (collect), or this CTA's groups of kSets 128-row tiles (act)
    int64_t n_iter;
    if (MODE == MODE_COLLECT) n_iter = (int64_t)buf.T + 1;
    else {
        const int64_t n_groups = (act.n + kRows * C::kSets - 1) / (kRows * C::kSets);
        n_iter = (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x;
    }

    // ============================================ MMA-issuing warps (one per set) ===================================================
    if (warp_u >= (uint32_t)C::kCW) {
        const int set = (int)warp_u - C::kCW;
        uint64_t* bars = bars_all + set * kBars;
        const uint32_t S = smem_u32(sm);
        const uint32_t sPhi = S + (uint32_t)set * C::kSetBytes, sPlo = sPhi + C::kTile, sX = S + C::oX + 32u * (uint32_t)set;
        const uint32_t sW2hi = S + C::oW2hi, sW2lo = S + C::oW2lo, sWS = S + C::oWS;
        const uint32_t TT = tmem + (uint32_t)set * C::kSetCols;
        constexpr int KS = H / 16;
        auto kfeat = [](int kk) -> uint32_t { return (uint32_t)(kk >> 2) * kSlab + (uint32_t)(kk & 3) * 32u; };
        auto kw2 = [](int kk) -> uint32_t { return (uint32_t)(kk >> 2) * C::kW2Slab + (uint32_t)(kk & 3) * 32u; };
#pragma unroll 1
        for (int64_t it = 0; it < n_iter; ++it) {
            const uint32_t p = (uint32_t)it & 1u;
            mbar_wait(&bars[RDY_X], p); fence_after_sync();
            if (elect_one()) {
                mma_f16(TT + C::cAcc, desc(sX), desc(sWS + 32u), idesc_f16(128, H, 0, 0), 0u);
                mma_f16(TT + C::cAcc, desc(sX), desc(sWS), idesc_f16(128, H, 0, 0), 1u);
                mma_commit(&bars[BAR_Z1]);
            }
            __syncwarp();
            mbar_wait(&bars[RDY_H1], p); fence_after_sync();
            if (elect_one()) {
                mma_f16(TT + C::cAcc, desc(sX), desc(sWS + 64u), idesc_f16(128, H, 0, 0), 0u);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a0 = pass == 0 ? sPlo : sPhi, b0 = pass == 1 ? sW2lo : sW2hi;
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) mma_f16(TT + C::cAcc, desc(a0 + kfeat(kk)), desc(b0 + kw2(kk)), idesc_f16(128, H, 0, 0), 1u);
                }
                mma_commit(&bars[BAR_Z2]);
            }
            __syncwarp();
            mbar_wait(&bars[RDY_H2], p); fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const uint32_t a0 = pass == 0 ? sPlo : sPhi;
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk)
                        mma_f16(TT + C::cH, desc(a0 + kfeat(kk)), desc(sWS + 96u + (uint32_t)kk * 2048u, C::kW2Slab), idesc_f16(128, 16, 0, 1), (pass | kk) ? 1u : 0u);
                }
                mma_commit(&bars[BAR_OUT]);
            }
            __syncwarp();
        }
        return;
    }

    // ============================================ compute warps ==================================================================
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const bool env_thread = cg < C::kSets;                    // the warps of column group s own the environments / rows of set s
    const int my_set = env_thread ? cg : 0;
    const int sw = row & 7;
    const uint32_t my_off = (uint32_t)(cg >> 1) * kSlab + (uint32_t)row * 128u;
    const int c0 = (cg & 1) * 4;
    const uint32_t T = tmem + ((uint32_t)(quad * 32) << 16);
    unsigned char* Xrow = sm + C::oX + row * 128;
    uint64_t* my_bars = bars_all + my_set * kBars;

    // accumulator of set s -> activation -> (hi, lo) -> the set's operand tile
    auto stage = [&](int s) {
        float z[32];
        tmem_ld32(T + (uint32_t)s * C::kSetCols + C::cAcc + 32 * cg, z);
        tmem_ld_wait();
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) split_pair(act_fwd(z[2 * e], m.act), act_fwd(z[2 * e + 1], m.act), hw[e], lw[e]);
        unsigned char* P = sm + (size_t)s * C::kSetBytes;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t o = my_off + (uint32_t)(((c0 + c) ^ sw) << 4);
            *reinterpret_cast<uint4*>(P + o) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
            *reinterpret_cast<uint4*>(P + C::kTile + o) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
        }
    };
    auto warp_ready = [&](uint64_t* bar) {
        fence_proxy_async();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    };
    // x16 = [x_hi | x_lo | 1 1] of one observation -> 16-column group my_set of the row (zeros for rows outside the batch)
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
        *reinterpret_cast<uint4*>(Xrow + (((2 * my_set) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(Xrow + (((2 * my_set + 1) ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
    };

    // ---- env-thread state ---------------------------------------------------------------------------------------------------
    const int64_t n_env = MODE == MODE_COLLECT ? ((int64_t)blockIdx.x * C::kSets + my_set) * kRows + row : 0;
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

    // the forward pass of iteration `it` starts from this row's observation ...
    auto produce = [&](int64_t it) {
        if (MODE == MODE_COLLECT) put_x16(o, D, owner);
        else {
            arow = (((int64_t)blockIdx.x + it * gridDim.x) * C::kSets + my_set) * kRows + row;
            aok = arow < act.n;
#pragma unroll
            for (int d = 0; d < D; ++d) o[d] = (aok && d < m.D) ? __ldg(act.obs + arow * m.D + d) : 0.f;
            put_x16(o, m.D, aok);
        }
        warp_ready(&my_bars[RDY_X]);                          // (its tcgen05 fence also orders the read of the previous head outputs)
    };
    // ... and ends with its head outputs: draw the action, write the step, advance the environment
    auto consume = [&](int64_t it) {
        mbar_wait(&my_bars[BAR_OUT], (uint32_t)it & 1u);
        fence_after_sync();
        float c[16];
        tmem_ld16(T + (uint32_t)my_set * C::kSetCols + C::cH, c);
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
#pragma unroll
        for (int s = 0; s < C::kSets; ++s) {
            mbar_wait(&bars_all[s * kBars + BAR_Z1], p);
            fence_after_sync();
            stage(s);
            warp_ready(&bars_all[s * kBars + RDY_H1]);
        }
#pragma unroll
        for (int s = 0; s < C::kSets; ++s) {
            mbar_wait(&bars_all[s * kBars + BAR_Z2], p);
            fence_after_sync();
            stage(s);
            warp_ready(&bars_all[s * kBars + RDY_H2]);
        }
    }
    if (env_thread && n_iter > 0) consume(n_iter - 1);
    fence_before_sync();
    asm volatile("bar.sync 1, %0;" ::"n"(C::kCompute) : "memory");
    if (warp == 0) tmem_dealloc(tmem, C::kCols);
}

// ---- host launchers (called from rollout_kernels.cu) --------------------------------------------------------------------------
bool f16_rollout_path(const gs_mlp_t* m) {
    const char* e = getenv("GS_ROLLOUT_IMPL");
    if (e && strcmp(e, "simt") == 0) return false;
    return m->hidden1 == m->hidden2 && (m->hidden1 == 64 || m->hidden1 == 128 || m->hidden1 == 256) && m->obs_dim <= hfu::kMaxD;
}
// 256 x 256: collect_wide.cu (W2 streamed)
int launch_collect_wide(gs_env* env, const MlpDev& md, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0, int deterministic,
                        cudaStream_t st);
int launch_policy_act_wide(const MlpDev& md, const float* obs, int64_t n, uint64_t seed, uint64_t offset, int64_t row_offset, int deterministic,
                           const float* uniforms, int32_t* actions, float* logp, float* value, float* logits, cudaStream_t st);

template <int H, int KIND>
static int launch_collect_hk(gs_env* env, const MlpDev& md, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0, int deterministic,
                             cudaStream_t st) {
    using C = cfu::Cfg<H>;
    auto kern = collect_f16_kernel<H, KIND, cfu::MODE_COLLECT>;
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    const int64_t per_cta = (int64_t)hfu::kRows * C::kSets;
    const unsigned grid = (unsigned)((env->n + per_cta - 1) / per_cta);
    kern<<<grid, C::kThreads, C::kSmemBytes, st>>>(to_dev(env), md, buf, cur_obs, cfu::ActDev{}, seed, step0, deterministic);
    GS_LAUNCH_CHECK();
    return 0;
}

int launch_collect_f16(gs_env* env, const MlpDev& md, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0, int deterministic,
                       cudaStream_t st) {
#define GS_CF(H, K) return launch_collect_hk<H, K>(env, md, buf, cur_obs, seed, step0, deterministic, st)
    if (md.H1 == 256) return launch_collect_wide(env, md, buf, cur_obs, seed, step0, deterministic, st);
    if (md.H1 == 64) {
        if (env->kind == GS_ENV_CARTPOLE_V1) GS_CF(64, GS_ENV_CARTPOLE_V1);
        if (env->kind == GS_ENV_ACROBOT_V1) GS_CF(64, GS_ENV_ACROBOT_V1);
        GS_CF(64, GS_ENV_MOUNTAINCAR_V0);
    }
    if (env->kind == GS_ENV_CARTPOLE_V1) GS_CF(128, GS_ENV_CARTPOLE_V1);
    if (env->kind == GS_ENV_ACROBOT_V1) GS_CF(128, GS_ENV_ACROBOT_V1);
    GS_CF(128, GS_ENV_MOUNTAINCAR_V0);
#undef GS_CF
}

template <int H>
static int launch_act_h(const MlpDev& md, const cfu::ActDev& act, uint64_t seed, uint64_t offset, int deterministic, int device, cudaStream_t st) {
    using C = cfu::Cfg<H>;
    auto kern = collect_f16_kernel<H, GS_ENV_CARTPOLE_V1, cfu::MODE_ACT>;
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    const int64_t per_group = (int64_t)hfu::kRows * C::kSets;
    const int64_t n_groups = (act.n + per_group - 1) / per_group;
    const int64_t cap = (int64_t)C::kMinCtas * sm_count(device);
    const unsigned grid = (unsigned)(n_groups < cap ? n_groups : cap);
    kern<<<grid, C::kThreads, C::kSmemBytes, st>>>(EnvDev{}, md, RolloutDev{}, nullptr, act, seed, offset, deterministic);
    GS_LAUNCH_CHECK();
    return 0;
}

int launch_policy_act_f16(const MlpDev& md, const float* obs, int64_t n, uint64_t seed, uint64_t offset, int64_t row_offset, int deterministic,
                          const float* uniforms, int32_t* actions, float* logp, float* value, float* logits, cudaStream_t st) {
    if (md.H1 == 256) return launch_policy_act_wide(md, obs, n, seed, offset, row_offset, deterministic, uniforms, actions, logp, value, logits, st);
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    cfu::ActDev act;
    act.obs = obs; act.n = n; act.row_offset = row_offset; act.uniforms = uniforms; act.actions = actions; act.logp = logp; act.value = value;
    act.logits_out = logits;
    if (md.H1 == 64) return launch_act_h<64>(md, act, seed, offset, deterministic, device, st);
    return launch_act_h<128>(md, act, seed, offset, deterministic, device, st);
}

}  // namespace gs
