// rollout_kernels.cu — policy_act / policy_predict_values and the fused persistent collect kernel.
//
// Replaces utils/policy_ops.py:14-41 (policy_act, policy_predict_values), the per-step body of
// RolloutCollector._collect (utils/rollout_collector.py:474-542: obs -> torch forward -> Categorical sample/log_prob ->
// env.step -> RolloutBuffer.add) and RolloutBuffer.add itself (utils/rollout_buffer.py:82-102).
//
// Collect is ONE launch per rollout: a CTA owns S environments for all T steps, env state lives in registers of the
// S "env threads", all 256 threads of the CTA run the MLP tile GEMMs between env steps, and each step's
// obs/action/logp/value/reward/done/timeout go straight into the time-major buffer as fully coalesced stores.
#include "rollout_shared.cuh"

namespace gs {

__device__ __forceinline__ void write_xs_row(float* sm_xs, int row, const float* o, int D) {
    float x[kDP];
#pragma unroll
    for (int d = 0; d < kDP; ++d) x[d] = d < D ? o[d] : 0.f;
    float4* xr = reinterpret_cast<float4*>(sm_xs + row * kLDX);
    xr[0] = make_float4(x[0], x[1], x[2], x[3]);
    xr[1] = make_float4(x[4], x[5], x[6], x[7]);
}

// ---- standalone policy_act over an (n, D) observation array ----------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(kThreads)
policy_act_kernel(MlpDev m, const float* __restrict__ obs, int64_t n, uint64_t rng_seed, uint64_t rng_offset, int64_t row_offset,
                  int deterministic, const float* __restrict__ uniforms, int32_t* __restrict__ actions, float* __restrict__ logp,
                  float* __restrict__ value, float* __restrict__ logits_out) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    load_resident<C>(sm, m);
    __syncthreads();
    const int64_t n_tiles = (n + C::S - 1) / C::S;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * C::S + tid;
        if (tid < C::S) {
            float o[kDP];
#pragma unroll
            for (int d = 0; d < kDP; ++d) o[d] = (row < n && d < m.D) ? __ldg(obs + row * m.D + d) : 0.f;
            write_xs_row(sm + C::oXS, tid, o, kDP);
        }
        __syncthreads();
        forward_backbone<C, false>(sm, m, C::S, nullptr);
        if (tid < C::S && row < n) {
            float out[kNH];
            forward_heads<C>(sm, tid, out);
            if (actions) {
                const float u = deterministic ? 0.f : (uniforms ? __ldg(uniforms + row) : action_uniform(rng_seed, (uint64_t)(row_offset + row), rng_offset));
                int a; float lp, v;
                act_from_heads(out, m.A, m.has_value, deterministic != 0, u, a, lp, v);
                actions[row] = a;
                if (logp) logp[row] = lp;
                if (value) value[row] = v;
            } else if (value) {
                value[row] = m.has_value ? (m.A == 2 ? out[2] : out[3]) : 0.f;
            }
            if (logits_out)
                for (int k = 0; k < m.A; ++k) logits_out[row * m.A + k] = k == 0 ? out[0] : (k == 1 ? out[1] : out[2]);
        }
        __syncthreads();
    }
}

// ---- fused collect ----------------------------------------------------------------------------------------------------
#ifndef GS_COLLECT_MIN_CTAS
#define GS_COLLECT_MIN_CTAS 3
#endif
template <class C, int KIND>
__global__ void __launch_bounds__(kThreads, (C::H1 <= 64 ? GS_COLLECT_MIN_CTAS : 1))   // 64-wide nets: 3 CTAs per SM (59.6 KB shared memory each) need <= 85 registers
collect_kernel(EnvDev h, MlpDev m, RolloutDev buf, float* __restrict__ cur_obs, uint64_t rng_seed, uint64_t step0, int deterministic) {
    extern __shared__ __align__(16) float sm[];
    constexpr int D = EnvDims<KIND>::D;
    const int tid = threadIdx.x;
    const int64_t n = (int64_t)blockIdx.x * C::S + tid;
    const bool owner = tid < C::S && n < h.n;
    load_resident<C>(sm, m);

    EnvRegs e;
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    if (owner) {
        env_load<KIND>(h, n, e);
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = cur_obs[n * D + d];
    }
    const uint64_t gid = (uint64_t)(h.params.gid0 + n);

    for (int t = 0; t < buf.T; ++t) {
        if (tid < C::S) write_xs_row(sm + C::oXS, tid, o, D);
        __syncthreads();
        forward_backbone<C, false>(sm, m, C::S, nullptr);
        if (owner) {
            float out[kNH];
            forward_heads<C>(sm, tid, out);
            const float u = deterministic ? 0.f : action_uniform(rng_seed, gid, step0 + (uint64_t)t);
            int a; float lp, v;
            act_from_heads(out, m.A, m.has_value, deterministic != 0, u, a, lp, v);
            const int64_t off = (int64_t)t * buf.N + n;
            store_obs_row<D>(buf.obs, off, o);
            buf.actions[off] = a;
            buf.logprobs[off] = lp;
            buf.values[off] = v;
            double r, ep_r;
            bool term, trunc;
            int ep_l;
            env_vec_step<KIND>(e, h.params, n, a, o, r, term, trunc, ep_r, ep_l);
            buf.rewards[off] = (float)r;
            buf.dones[off] = (term || trunc) ? 1 : 0;
            buf.timeouts[off] = trunc ? 1 : 0;
            if (buf.next_obs) store_obs_row<D>(buf.next_obs, off, o);
            if (buf.ep_return) buf.ep_return[off] = ep_r;
            if (buf.ep_length) buf.ep_length[off] = ep_l;
        }
        __syncthreads();
    }
    // V(last_obs) for the GAE bootstrap (rollout_collector.py:373) and the state hand-back
    if (tid < C::S) write_xs_row(sm + C::oXS, tid, o, D);
    __syncthreads();
    forward_backbone<C, false>(sm, m, C::S, nullptr);
    if (owner) {
        float out[kNH];
        forward_heads<C>(sm, tid, out);
        if (buf.last_values) buf.last_values[n] = m.has_value ? (m.A == 2 ? out[2] : out[3]) : 0.f;
        if (buf.last_obs) store_obs_row<D>(buf.last_obs, n, o);
        store_obs_row<D>(cur_obs, n, o);
        env_store<KIND>(h, n, e);
    }
}

int validate_mlp(const gs_mlp_t* m);  // update_kernels.cu
// collect_f16.cu: the same two entry points with the forward pass on the tensor cores (64x64 / 128x128 networks)
bool f16_rollout_path(const gs_mlp_t* m);
int launch_collect_f16(gs_env* env, const MlpDev& md, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0, int deterministic,
                       cudaStream_t st);
int launch_policy_act_f16(const MlpDev& md, const float* obs, int64_t n, uint64_t seed, uint64_t offset, int64_t row_offset, int deterministic,
                          const float* uniforms, int32_t* actions, float* logp, float* value, float* logits, cudaStream_t st);

static MlpDev mlp_dev(const gs_mlp_t* m) {
    MlpDev d;
    d.D = m->obs_dim; d.H1 = m->hidden1; d.H2 = m->hidden2; d.A = m->n_actions; d.has_value = m->has_value; d.act = m->activation;
    d.w1 = m->w1; d.b1 = m->b1; d.w2 = m->w2; d.b2 = m->b2; d.wp = m->wp; d.bp = m->bp; d.wv = m->wv; d.bv = m->bv;
    return d;
}

template <class C>
static int launch_policy_act(const gs_mlp_t* m, const float* obs, int64_t n, uint64_t seed, uint64_t offset, int64_t row_offset,
                             int deterministic, const float* uniforms, int32_t* actions, float* logp, float* value, float* logits,
                             cudaStream_t st) {
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    const int64_t n_tiles = (n + C::S - 1) / C::S;
    const int64_t cap = 4ll * sm_count(device);
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    const size_t smem = (size_t)C::kSmemFloats * sizeof(float);
    GS_CUDA(cudaFuncSetAttribute(policy_act_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_act_kernel<C><<<grid, kThreads, smem, st>>>(mlp_dev(m), obs, n, seed, offset, row_offset, deterministic, uniforms, actions,
                                                       logp, value, logits);
    GS_LAUNCH_CHECK();
    return 0;
}

static int dispatch_policy_act(const gs_mlp_t* m, const float* obs, int64_t n, uint64_t seed, uint64_t offset, int64_t row_offset,
                               int deterministic, const float* uniforms, int32_t* actions, float* logp, float* value, float* logits,
                               cudaStream_t st) {
    if (validate_mlp(m)) return -1;
    if (!obs || n <= 0) GS_FAIL("policy_act: empty observation batch");
    if (f16_rollout_path(m)) return launch_policy_act_f16(mlp_dev(m), obs, n, seed, offset, row_offset, deterministic, uniforms, actions, logp, value, logits, st);
    const int h1 = m->hidden1, h2 = m->hidden2;
#define GS_PA(H1, H2, S) return launch_policy_act<TileCfg<H1, H2, S>>(m, obs, n, seed, offset, row_offset, deterministic, uniforms, actions, logp, value, logits, st)
    if (h1 == 64 && h2 == 64) GS_PA(64, 64, 64);
    if (h1 == 64 && h2 == 0) GS_PA(64, 0, 64);
    if (h1 == 128 && h2 == 128) GS_PA(128, 128, 64);
    if (h1 == 256 && h2 == 256) GS_PA(256, 256, 64);
#undef GS_PA
    GS_FAIL("no kernel for hidden dims (%d,%d)", h1, h2);
}

template <class C, int KIND>
static int launch_collect(gs_env* env, const gs_mlp_t* m, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0,
                          int deterministic, cudaStream_t st) {
    const unsigned grid = (unsigned)((env->n + C::S - 1) / C::S);
    const size_t smem = (size_t)C::kSmemFloats * sizeof(float);
    GS_CUDA(cudaFuncSetAttribute(collect_kernel<C, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    collect_kernel<C, KIND><<<grid, kThreads, smem, st>>>(to_dev(env), mlp_dev(m), buf, cur_obs, seed, step0, deterministic);
    GS_LAUNCH_CHECK();
    return 0;
}

template <int KIND>
static int dispatch_collect(gs_env* env, const gs_mlp_t* m, const RolloutDev& buf, float* cur_obs, uint64_t seed, uint64_t step0,
                            int deterministic, cudaStream_t st) {
    const int h1 = m->hidden1, h2 = m->hidden2;
#define GS_CO(H1, H2, S) return launch_collect<TileCfg<H1, H2, S>, KIND>(env, m, buf, cur_obs, seed, step0, deterministic, st)
    if (h1 == 64 && h2 == 64) GS_CO(64, 64, 64);
    if (h1 == 64 && h2 == 0) GS_CO(64, 0, 64);
    if (h1 == 128 && h2 == 128) GS_CO(128, 128, 64);
    if (h1 == 256 && h2 == 256) GS_CO(256, 256, 64);
#undef GS_CO
    GS_FAIL("no kernel for hidden dims (%d,%d)", h1, h2);
}

}  // namespace gs

using namespace gs;

extern "C" {

int gs_policy_act(const gs_mlp_t* mlp, const float* obs, int64_t n, uint64_t rng_seed, uint64_t rng_offset, int64_t row_offset,
                  int deterministic, const float* uniforms, int32_t* actions, float* logp, float* value, float* logits_out,
                  void* stream) {
    if (!actions || !logp || !value) GS_FAIL("gs_policy_act: NULL output");
    return dispatch_policy_act(mlp, obs, n, rng_seed, rng_offset, row_offset, deterministic, uniforms, actions, logp, value, logits_out,
                               (cudaStream_t)stream);
}

int gs_policy_values(const gs_mlp_t* mlp, const float* obs, int64_t n, float* value, void* stream) {
    if (!value) GS_FAIL("gs_policy_values: NULL output");
    return dispatch_policy_act(mlp, obs, n, 0, 0, 0, 1, nullptr, nullptr, nullptr, value, nullptr, (cudaStream_t)stream);
}

int gs_rollout_collect(gs_env_t* env, const gs_mlp_t* mlp, const gs_rollout_t* b, float* cur_obs, uint64_t rng_seed,
                       uint64_t rng_offset, int deterministic, void* stream) {
    if (!env || !b || !cur_obs) GS_FAIL("gs_rollout_collect: NULL argument");
    if (validate_mlp(mlp)) return -1;
    if (b->N != env->n) GS_FAIL("gs_rollout_collect: buffer N=%lld != env n=%lld", (long long)b->N, (long long)env->n);
    if (b->T <= 0) GS_FAIL("gs_rollout_collect: n_steps must be > 0");
    if (b->obs_dim != gs_env_obs_dim(env->kind) || mlp->obs_dim != b->obs_dim) GS_FAIL("gs_rollout_collect: obs_dim mismatch");
    if (mlp->n_actions != gs_env_n_actions(env->kind)) GS_FAIL("gs_rollout_collect: policy has %d actions, env has %d", mlp->n_actions, gs_env_n_actions(env->kind));
    if (!b->obs || !b->actions || !b->logprobs || !b->values || !b->rewards || !b->dones || !b->timeouts)
        GS_FAIL("gs_rollout_collect: buffer has NULL arrays");
    RolloutDev d;
    d.T = b->T; d.D = b->obs_dim; d.N = b->N; d.obs = b->obs; d.next_obs = b->next_obs; d.actions = b->actions;
    d.logprobs = b->logprobs; d.values = b->values; d.rewards = b->rewards; d.dones = b->dones; d.timeouts = b->timeouts;
    d.last_obs = b->last_obs; d.last_values = b->last_values; d.ep_return = b->ep_return; d.ep_length = b->ep_length;
    cudaStream_t st = (cudaStream_t)stream;
    if (f16_rollout_path(mlp)) return launch_collect_f16(env, mlp_dev(mlp), d, cur_obs, rng_seed, rng_offset, deterministic, st);
    switch (env->kind) {
        case GS_ENV_CARTPOLE_V1: return dispatch_collect<GS_ENV_CARTPOLE_V1>(env, mlp, d, cur_obs, rng_seed, rng_offset, deterministic, st);
        case GS_ENV_ACROBOT_V1: return dispatch_collect<GS_ENV_ACROBOT_V1>(env, mlp, d, cur_obs, rng_seed, rng_offset, deterministic, st);
        default: return dispatch_collect<GS_ENV_MOUNTAINCAR_V0>(env, mlp, d, cur_obs, rng_seed, rng_offset, deterministic, st);
    }
}

}  // extern "C"
