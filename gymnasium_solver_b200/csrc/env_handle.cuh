// env_handle.cuh — the opaque gs_env_t: SoA device state of N environments.
#pragma once

#include "env_dynamics.cuh"

struct gs_env {
    int kind;
    int device;
    int64_t n;
    gs::EnvParams params;   // kind, max_steps, wrapper, wrapper params, counts table, seed, gid0
    // HBM-resident state, SoA
    double* state;          // [S][N]
    double* ep_ret;         // [N]   RecordEpisodeStatistics.episode_returns
    int32_t* elapsed;       // [N]   TimeLimit._elapsed_steps
    int32_t* ep_len;        // [N]   RecordEpisodeStatistics.episode_lengths
    uint32_t* reset_count;  // [N]   Philox counter of the reset stream
    uint8_t* needs_reset;   // [N]   SyncVectorEnv._autoreset_envs
};

namespace gs {

struct EnvDev {  // by-value kernel argument
    int64_t n;
    EnvParams params;
    double* state;
    double* ep_ret;
    int32_t* elapsed;
    int32_t* ep_len;
    uint32_t* reset_count;
    uint8_t* needs_reset;
};

inline EnvDev to_dev(const gs_env* e) {
    EnvDev d;
    d.n = e->n; d.params = e->params; d.state = e->state; d.ep_ret = e->ep_ret; d.elapsed = e->elapsed;
    d.ep_len = e->ep_len; d.reset_count = e->reset_count; d.needs_reset = e->needs_reset;
    return d;
}

template <int KIND>
__device__ __forceinline__ void env_load(const EnvDev& h, int64_t i, EnvRegs& e) {
#pragma unroll
    for (int k = 0; k < 4; ++k) e.s[k] = k < EnvDims<KIND>::S ? h.state[(int64_t)k * h.n + i] : 0.0;
    e.ep_ret = h.ep_ret[i];
    e.elapsed = h.elapsed[i];
    e.ep_len = h.ep_len[i];
    e.reset_count = h.reset_count[i];
    e.needs_reset = h.needs_reset[i];
}

template <int KIND>
__device__ __forceinline__ void env_store(const EnvDev& h, int64_t i, const EnvRegs& e) {
#pragma unroll
    for (int k = 0; k < EnvDims<KIND>::S; ++k) h.state[(int64_t)k * h.n + i] = e.s[k];
    h.ep_ret[i] = e.ep_ret;
    h.elapsed[i] = e.elapsed;
    h.ep_len[i] = e.ep_len;
    h.reset_count[i] = e.reset_count;
    h.needs_reset[i] = (uint8_t)e.needs_reset;
}

}  // namespace gs
