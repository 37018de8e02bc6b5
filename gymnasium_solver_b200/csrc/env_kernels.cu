// env_kernels.cu — GPU-resident vector environments behind the Gymnasium VectorEnv protocol
// (reset / step as consumed at utils/rollout_collector.py:317 and :504), one thread per env.
// HBM-bound: state is SoA fp64 ([S][N]) so every load / store of a warp is one contiguous line; the per-step
// algorithmic traffic is the 98 / 106 / 58 B per env-step of SURVEY.md §8(d).
#include <cmath>
#include <vector>

#include "env_handle.cuh"

namespace gs {

constexpr int kEnvThreads = 128;

template <int KIND>
__global__ void __launch_bounds__(kEnvThreads) env_reset_kernel(EnvDev h, float* __restrict__ obs) {
    const int64_t i = (int64_t)blockIdx.x * kEnvThreads + threadIdx.x;
    if (i >= h.n) return;
    EnvRegs e;
    e.reset_count = h.reset_count[i];
    env_reset_state<KIND>(e.s, h.params.seed, (uint64_t)(h.params.gid0 + i), e.reset_count);
    if (KIND == GS_ENV_MOUNTAINCAR_V0 && h.params.wrapper == GS_WRAP_SCRIPTED_REPLAY) {   // scripted replay: reset shows obs[0] of the table
        const float* tab = reinterpret_cast<const float*>(h.params.counts);
        const int L = (int)h.params.wp[0];
        e.s[0] = (double)tab[3 * L]; e.s[1] = (double)tab[3 * L + 1];
    }
    e.reset_count += 1;
    e.elapsed = 0; e.ep_ret = 0.0; e.ep_len = 0; e.needs_reset = 0;
    env_store<KIND>(h, i, e);
    float o[EnvDims<KIND>::D];
    env_obs<KIND>(e.s, o);
    obs_normalize<EnvDims<KIND>::D>(h.params, o);
#pragma unroll
    for (int d = 0; d < EnvDims<KIND>::D; ++d) obs[i * EnvDims<KIND>::D + d] = o[d];
}

template <int KIND>
__global__ void __launch_bounds__(kEnvThreads)
env_step_kernel(EnvDev h, const int32_t* __restrict__ actions, float* __restrict__ obs, float* __restrict__ reward,
                uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated, double* __restrict__ ep_return,
                int32_t* __restrict__ ep_length) {
    const int64_t i = (int64_t)blockIdx.x * kEnvThreads + threadIdx.x;
    if (i >= h.n) return;
    EnvRegs e;
    // the reset-stream counter is only touched on the (rare) autoreset step: 8 B of the per-step HBM traffic saved
#pragma unroll
    for (int k = 0; k < 4; ++k) e.s[k] = k < EnvDims<KIND>::S ? h.state[(int64_t)k * h.n + i] : 0.0;
    e.ep_ret = h.ep_ret[i];
    e.elapsed = h.elapsed[i];
    e.ep_len = h.ep_len[i];
    e.needs_reset = h.needs_reset[i];
    const bool resetting = e.needs_reset != 0;
    e.reset_count = resetting ? h.reset_count[i] : 0u;
    float o[EnvDims<KIND>::D];
    double r, ep_r;
    bool term, trunc;
    int ep_l;
    env_vec_step<KIND>(e, h.params, i, __ldg(actions + i), o, r, term, trunc, ep_r, ep_l);
#pragma unroll
    for (int k = 0; k < EnvDims<KIND>::S; ++k) h.state[(int64_t)k * h.n + i] = e.s[k];
    h.ep_ret[i] = e.ep_ret;
    h.elapsed[i] = e.elapsed;
    h.ep_len[i] = e.ep_len;
    if (resetting) h.reset_count[i] = e.reset_count;
    h.needs_reset[i] = (uint8_t)e.needs_reset;
    if (EnvDims<KIND>::D == 4) {
        reinterpret_cast<float4*>(obs)[i] = make_float4(o[0], o[1], o[2], o[3]);
    } else if (EnvDims<KIND>::D == 2) {
        reinterpret_cast<float2*>(obs)[i] = make_float2(o[0], o[1]);
    } else {
#pragma unroll
        for (int d = 0; d < EnvDims<KIND>::D; ++d) obs[i * EnvDims<KIND>::D + d] = o[d];
    }
    reward[i] = (float)r;
    terminated[i] = term ? 1 : 0;
    truncated[i] = trunc ? 1 : 0;
    if (ep_return) ep_return[i] = ep_r;
    if (ep_length) ep_length[i] = ep_l;
}

static int state_dim(int kind) { return kind == GS_ENV_MOUNTAINCAR_V0 ? 2 : 4; }

}  // namespace gs

using namespace gs;

extern "C" {

int gs_env_obs_dim(int kind) { return kind == GS_ENV_CARTPOLE_V1 ? 4 : (kind == GS_ENV_ACROBOT_V1 ? 6 : (kind == GS_ENV_MOUNTAINCAR_V0 ? 2 : -1)); }
int gs_env_state_dim(int kind) { return (kind < 0 || kind > 2) ? -1 : state_dim(kind); }
int gs_env_n_actions(int kind) { return kind == GS_ENV_CARTPOLE_V1 ? 2 : ((kind == GS_ENV_ACROBOT_V1 || kind == GS_ENV_MOUNTAINCAR_V0) ? 3 : -1); }
int64_t gs_env_num_envs(const gs_env_t* env) { return env ? env->n : -1; }

int gs_env_create(int kind, int64_t n_envs, int64_t env_id_offset, uint64_t seed, int max_episode_steps, int device, gs_env_t** out) {
    if (!out) GS_FAIL("gs_env_create: out is NULL");
    if (kind < 0 || kind > 2) GS_FAIL("gs_env_create: unknown env kind %d", kind);
    if (n_envs <= 0) GS_FAIL("gs_env_create: n_envs must be > 0");
    if (max_episode_steps < 0) GS_FAIL("gs_env_create: max_episode_steps must be >= 0");
    GS_CUDA(cudaSetDevice(device));
    gs_env* e = new gs_env();
    memset(e, 0, sizeof(gs_env));
    e->kind = kind; e->device = device; e->n = n_envs;
    e->params.kind = kind;
    e->params.max_steps = max_episode_steps > 0 ? max_episode_steps : (kind == GS_ENV_MOUNTAINCAR_V0 ? 200 : 500);
    e->params.wrapper = 0;
    e->params.counts = nullptr;
    e->params.obs_norm = 0;
    for (int d = 0; d < 6; ++d) { e->params.on_mode[d] = 0; e->params.on_low[d] = 0.f; e->params.on_den[d] = 1.f; }
    e->params.seed = seed;
    e->params.gid0 = env_id_offset;
    const size_t n = (size_t)n_envs;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) { err = cudaMalloc(p, bytes); if (err == cudaSuccess) err = cudaMemset(*p, 0, bytes); } };
    alloc((void**)&e->state, sizeof(double) * state_dim(kind) * n);
    alloc((void**)&e->ep_ret, sizeof(double) * n);
    alloc((void**)&e->elapsed, sizeof(int32_t) * n);
    alloc((void**)&e->ep_len, sizeof(int32_t) * n);
    alloc((void**)&e->reset_count, sizeof(uint32_t) * n);
    alloc((void**)&e->needs_reset, n);
    if (err != cudaSuccess) {
        gs_env_destroy(e);
        GS_FAIL("gs_env_create: allocation failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return 0;
}

int gs_env_destroy(gs_env_t* e) {
    if (!e) return 0;
    cudaFree(e->state); cudaFree(e->ep_ret); cudaFree(e->elapsed); cudaFree(e->ep_len);
    cudaFree(e->reset_count); cudaFree(e->needs_reset); cudaFree(e->params.counts);
    delete e;
    return 0;
}

int gs_wrapper_attach(gs_env_t* e, int wrapper_kind, const double* params_host, int n_params) {
    if (!e || !params_host) GS_FAIL("gs_wrapper_attach: NULL argument");
    if (e->params.wrapper) GS_FAIL("gs_wrapper_attach: env already has wrapper %d (one device wrapper per env)", e->params.wrapper);
    int need = 0, kind = -1;
    switch (wrapper_kind) {
        case GS_WRAP_MOUNTAINCAR_STATE_COUNT_BONUS: need = 5; kind = GS_ENV_MOUNTAINCAR_V0; break;
        case GS_WRAP_CARTPOLE_REWARD_SHAPER: need = 3; kind = GS_ENV_CARTPOLE_V1; break;
        case GS_WRAP_MOUNTAINCAR_REWARD_SHAPER: need = 3; kind = GS_ENV_MOUNTAINCAR_V0; break;
        case GS_WRAP_SCRIPTED_REPLAY: {
            kind = GS_ENV_MOUNTAINCAR_V0;
            const int L = n_params > 0 ? (int)params_host[0] : 0;
            if (L < 1 || L > 65536) GS_FAIL("gs_wrapper_attach: scripted replay needs 1..65536 steps, got %d", L);
            need = 1 + 3 * L + 2 * (L + 1);
            break;
        }
        default: GS_FAIL("gs_wrapper_attach: unknown wrapper kind %d", wrapper_kind);
    }
    if (e->kind != kind) GS_FAIL("gs_wrapper_attach: wrapper %d does not apply to env kind %d", wrapper_kind, e->kind);
    if (n_params != need) GS_FAIL("gs_wrapper_attach: wrapper %d takes %d params, got %d", wrapper_kind, need, n_params);
    for (int i = 0; i < 5; ++i) e->params.wp[i] = i < n_params ? params_host[i] : 0.0;
    if (wrapper_kind == GS_WRAP_SCRIPTED_REPLAY) {            // the tables live on the device as floats behind params.counts
        std::vector<float> tab(n_params - 1);
        for (int i = 1; i < n_params; ++i) tab[i - 1] = (float)params_host[i];
        GS_CUDA(cudaSetDevice(e->device));
        GS_CUDA(cudaMalloc((void**)&e->params.counts, sizeof(float) * tab.size()));
        GS_CUDA(cudaMemcpy(e->params.counts, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice));
    }
    if (wrapper_kind == GS_WRAP_MOUNTAINCAR_STATE_COUNT_BONUS) {
        const double pb = params_host[0], vb = params_host[1];
        if (pb < 1 || vb < 1 || pb > 4096 || vb > 4096) GS_FAIL("gs_wrapper_attach: bins out of range");
        if (params_host[3] < 0 || params_host[3] > 2) GS_FAIL("Unknown bonus_type: %g", params_host[3]);
        const size_t bytes = sizeof(uint32_t) * (size_t)pb * (size_t)vb * (size_t)e->n;
        GS_CUDA(cudaSetDevice(e->device));
        GS_CUDA(cudaMalloc((void**)&e->params.counts, bytes));
        GS_CUDA(cudaMemset(e->params.counts, 0, bytes));
    }
    e->params.wrapper = wrapper_kind;
    return 0;
}

int gs_env_set_obs_normalization(gs_env_t* e, const float* low_host, const float* high_host, int n_dims) {
    if (!e) GS_FAIL("gs_env_set_obs_normalization: NULL env");
    const int D = gs_env_obs_dim(e->kind);
    if (!low_host || !high_host) { e->params.obs_norm = 0; return 0; }          // NULL bounds: switch it off
    if (n_dims != D) GS_FAIL("gs_env_set_obs_normalization: env has %d observation dims, got %d bounds", D, n_dims);
    for (int d = 0; d < D; ++d) {
        const float lo = low_host[d], hi = high_host[d];
        const bool finite = std::isfinite(lo) && std::isfinite(hi);
        e->params.on_mode[d] = (finite && hi > lo) ? 1 : ((finite && hi == lo) ? 2 : 0);
        e->params.on_low[d] = lo;
        const float scale = e->params.on_mode[d] == 1 ? hi - lo : 1.0f;         // vec_normalize_static.py:33: where(pos_scale, high - low, 1.0)
        e->params.on_den[d] = (float)((double)scale + 1e-8);                    // float32 array + Python float -> float32
    }
    e->params.obs_norm = 1;
    return 0;
}

int gs_env_set_state(gs_env_t* e, const double* state, const int32_t* elapsed, void* stream) {
    if (!e || !state) GS_FAIL("gs_env_set_state: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GS_CUDA(cudaMemcpyAsync(e->state, state, sizeof(double) * state_dim(e->kind) * (size_t)e->n, cudaMemcpyDeviceToDevice, st));
    if (elapsed) GS_CUDA(cudaMemcpyAsync(e->elapsed, elapsed, sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int gs_env_get_state(gs_env_t* e, double* state, int32_t* elapsed, void* stream) {
    if (!e || !state) GS_FAIL("gs_env_get_state: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GS_CUDA(cudaMemcpyAsync(state, e->state, sizeof(double) * state_dim(e->kind) * (size_t)e->n, cudaMemcpyDeviceToDevice, st));
    if (elapsed) GS_CUDA(cudaMemcpyAsync(elapsed, e->elapsed, sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---- exact snapshot of everything a handle holds on the device (checkpoint / resume) -------------------------------------
// blob layout (8-byte aligned sections): state [S][N] f64 | ep_ret [N] f64 | elapsed [N] i32 | ep_len [N] i32 |
// reset_count [N] u32 | needs_reset [N] u8 | StateCountBonus tables [N][pb*vb] u32 (when attached)
static size_t pad8(size_t b) { return (b + 7) & ~(size_t)7; }
static size_t counts_bytes(const gs_env* e) {
    if (e->params.wrapper != GS_WRAP_MOUNTAINCAR_STATE_COUNT_BONUS || !e->params.counts) return 0;
    return sizeof(uint32_t) * (size_t)e->params.wp[0] * (size_t)e->params.wp[1] * (size_t)e->n;
}
int64_t gs_env_snapshot_bytes(const gs_env_t* e) {
    if (!e) GS_FAIL("gs_env_snapshot_bytes: NULL argument");
    const size_t n = (size_t)e->n;
    return (int64_t)(pad8(sizeof(double) * state_dim(e->kind) * n) + pad8(sizeof(double) * n) + 3 * pad8(4 * n) + pad8(n) + pad8(counts_bytes(e)));
}
static int snapshot_copy(gs_env* e, char* blob, bool save, cudaStream_t st) {
    const size_t n = (size_t)e->n;
    struct Sec { void* p; size_t bytes; } secs[] = {
        {e->state, sizeof(double) * state_dim(e->kind) * n}, {e->ep_ret, sizeof(double) * n}, {e->elapsed, 4 * n}, {e->ep_len, 4 * n},
        {e->reset_count, 4 * n}, {e->needs_reset, n}, {e->params.counts, counts_bytes(e)}};
    for (const Sec& s : secs) {
        if (s.bytes) {
            if (save) GS_CUDA(cudaMemcpyAsync(blob, s.p, s.bytes, cudaMemcpyDeviceToDevice, st));
            else GS_CUDA(cudaMemcpyAsync(s.p, blob, s.bytes, cudaMemcpyDeviceToDevice, st));
        }
        blob += pad8(s.bytes);
    }
    return 0;
}
int gs_env_save(gs_env_t* e, void* blob, void* stream) {
    if (!e || !blob) GS_FAIL("gs_env_save: NULL argument");
    return snapshot_copy(e, (char*)blob, true, (cudaStream_t)stream);
}
int gs_env_load(gs_env_t* e, const void* blob, void* stream) {
    if (!e || !blob) GS_FAIL("gs_env_load: NULL argument");
    return snapshot_copy(e, (char*)blob, false, (cudaStream_t)stream);
}

int gs_env_reset(gs_env_t* e, float* obs, void* stream) {
    if (!e || !obs) GS_FAIL("gs_env_reset: NULL argument");
    const unsigned blocks = (unsigned)((e->n + kEnvThreads - 1) / kEnvThreads);
    cudaStream_t st = (cudaStream_t)stream;
    const EnvDev h = to_dev(e);
    switch (e->kind) {
        case GS_ENV_CARTPOLE_V1: env_reset_kernel<GS_ENV_CARTPOLE_V1><<<blocks, kEnvThreads, 0, st>>>(h, obs); break;
        case GS_ENV_ACROBOT_V1: env_reset_kernel<GS_ENV_ACROBOT_V1><<<blocks, kEnvThreads, 0, st>>>(h, obs); break;
        default: env_reset_kernel<GS_ENV_MOUNTAINCAR_V0><<<blocks, kEnvThreads, 0, st>>>(h, obs); break;
    }
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_env_step(gs_env_t* e, const int32_t* actions, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated,
                double* ep_return, int32_t* ep_length, void* stream) {
    if (!e || !actions || !obs || !reward || !terminated || !truncated) GS_FAIL("gs_env_step: NULL argument");
    const unsigned blocks = (unsigned)((e->n + kEnvThreads - 1) / kEnvThreads);
    cudaStream_t st = (cudaStream_t)stream;
    const EnvDev h = to_dev(e);
    switch (e->kind) {
        case GS_ENV_CARTPOLE_V1:
            env_step_kernel<GS_ENV_CARTPOLE_V1><<<blocks, kEnvThreads, 0, st>>>(h, actions, obs, reward, terminated, truncated, ep_return, ep_length); break;
        case GS_ENV_ACROBOT_V1:
            env_step_kernel<GS_ENV_ACROBOT_V1><<<blocks, kEnvThreads, 0, st>>>(h, actions, obs, reward, terminated, truncated, ep_return, ep_length); break;
        default:
            env_step_kernel<GS_ENV_MOUNTAINCAR_V0><<<blocks, kEnvThreads, 0, st>>>(h, actions, obs, reward, terminated, truncated, ep_return, ep_length); break;
    }
    GS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
