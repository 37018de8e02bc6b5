// rollout_shared.cuh — what the rollout kernels (rollout_kernels.cu: fp32 FMA pipe; collect_f16.cu: fp16x3 tensor cores) share:
// the device view of the rollout buffer, the action draw from the head outputs, observation-row stores.
#pragma once

#include "env_handle.cuh"
#include "mlp_tile.cuh"

namespace gs {

struct RolloutDev {
    int T, D;
    int64_t N;
    float *obs, *next_obs;
    int32_t* actions;
    float *logprobs, *values, *rewards;
    uint8_t *dones, *timeouts;
    float *last_obs, *last_values;
    double* ep_return;
    int32_t* ep_length;
};

// sample / mode + log-prob from the 4 head outputs of one row
__device__ __forceinline__ void act_from_heads(const float (&out)[kNH], int A, int has_value, bool deterministic, float u,
                                               int& action, float& logp, float& value) {
    float lp[3] = {0.f, 0.f, 0.f};
    log_softmax(out, A, lp);
    int a = 0;
    if (deterministic) {  // dist.mode: first maximum
        float best = lp[0];
#pragma unroll
        for (int k = 1; k < 3; ++k)
            if (k < A && lp[k] > best) { best = lp[k]; a = k; }
    } else {              // inverse CDF: a = #{k : cdf_k <= u}, clamped
        float cdf = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < A) { cdf += expf(lp[k]); a += (cdf <= u) ? 1 : 0; }
        a = a < A - 1 ? a : A - 1;
    }
    action = a;
    logp = a == 0 ? lp[0] : (a == 1 ? lp[1] : lp[2]);
    value = has_value ? (A == 2 ? out[2] : out[3]) : 0.0f;
}

template <int D>
__device__ __forceinline__ void store_obs_row(float* dst, int64_t row, const float* o) {
    if (D == 4) reinterpret_cast<float4*>(dst)[row] = make_float4(o[0], o[1], o[2], o[3]);
    else if (D == 2) reinterpret_cast<float2*>(dst)[row] = make_float2(o[0], o[1]);
    else {
#pragma unroll
        for (int d = 0; d < D; ++d) dst[row * D + d] = o[d];
    }
}

}  // namespace gs
