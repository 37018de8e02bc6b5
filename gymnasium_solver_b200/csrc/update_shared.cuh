// update_shared.cuh — definitions shared by the SIMT (update_kernels.cu) and tensor-core (update_tc.cu) update kernels:
// minibatch / hyper-parameter device structs, the per-CTA metric partial layout, sample-id translation and the
// per-sample loss + d(loss)/d(head outputs) arithmetic of PPOAgent / REINFORCEAgent.losses_for_batch.
#pragma once

#include "mlp_tile.cuh"

namespace gs {

enum { ALGO_PPO = 0, ALGO_REINFORCE = 1 };

// per-CTA partial sums (doubles)
enum {
    PM_SURR = 0, PM_VLOSS, PM_ENT, PM_CLIPF, PM_CLIPF_VF, PM_RV, PM_RV2, PM_R, PM_R2, PM_KL, PM_AKL, PM_ADVN, PM_ADVN2,
    PM_TGT, PM_TGT2, PM_RETN, PM_RETN2, PM_Z0, PM_Z0SQ, PM_Z1, PM_Z1SQ, PM_COUNT, PM_N
};

struct BatchDev {
    int64_t n;
    const int64_t* idx;
    uint64_t perm_key;
    int64_t perm_offset, perm_len;
    const int64_t* idx_map;
    int T, D;
    int64_t N;
    const float* obs;
    const int32_t* actions;
    const float *logp_old, *values_old, *adv, *ret;
    const float* packed;   // nullable: (T*N, 16) sample records (gs_rollout_pack)
    int prepared;          // gs_batch_prepare already wrote the sample offsets of this minibatch into the workspace
    int defer_reduce;      // leave the per-CTA partials in the workspace: gs_update_finish completes the step
    uint32_t* offsets;     // nullable: caller-owned offsets buffer (instead of the workspace's)
    int perm_bits;         // feistel_bits(perm_len), computed once on the host (0: unknown)
};

struct HpDev {
    float clip_lo, clip_hi, clip_vf, vf_coef, ent_coef;
    int normalize_adv, normalize_ret, policy_targets;
};

__device__ __forceinline__ int64_t sample_offset(const BatchDev& b, int64_t pos) {
    int64_t i;
    if (b.idx) i = b.idx[pos];
    else if (b.perm_len > 0) i = (int64_t)feistel_permute((uint64_t)(b.perm_offset + pos), (uint64_t)b.perm_len, b.perm_key, b.perm_bits);
    else i = b.perm_offset + pos;
    if (b.idx_map) i = b.idx_map[i];
    if ((uint64_t)i < (1ull << 32)) {              // 32-bit division (a 64-bit one by a runtime divisor is ~100 instructions)
        const uint32_t e = (uint32_t)i / (uint32_t)b.T, t = (uint32_t)i - e * (uint32_t)b.T;
        return (int64_t)t * b.N + e;
    }
    const int64_t e = i / b.T, t = i - e * b.T;   // env-major id -> (env, step)
    return t * b.N + e;                            // time-major offset
}

// mean / (std + eps) denominators from raw moments {sum, sumsq, count}; unbiased std like torch.std
__device__ __forceinline__ void norm_consts(const double* mom, float& mean, float& denom) {
    const double n = mom[2];
    const double mu = mom[0] / n;
    double var = (mom[1] - mom[0] * mu) / (n - 1.0);
    var = var > 0.0 ? var : 0.0;
    mean = (float)mu;
    denom = (float)sqrt(var) + 1e-8f;
}

// Per-sample loss terms, metric partial sums and the gradient w.r.t. the 4 head outputs (logits 0..A-1, value at index A).
// Reference arithmetic: agents/ppo/ppo_agent.py:37-129, agents/reinforce/reinforce_agent.py:23-72, utils/torch.py:102-119.
// torch.min / torch.max tie rule (gradient split evenly) and clamp's inclusive pass-through are reproduced.
template <int ALGO>
__device__ __forceinline__ void sample_loss(const float (&out)[kNH], int A, int a_s, float lp_old, float v_old, float adv_s, float ret_s,
                                            const HpDev& hp, float adv_mean, float adv_den, float ret_mean, float ret_den, float invB,
                                            float (&g)[kNH], float (&pm)[PM_N]) {
    float lp[3] = {0.f, 0.f, 0.f}, p[3] = {0.f, 0.f, 0.f};
    log_softmax(out, A, lp);
    float H = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (k < A) { p[k] = GS_EXPF(lp[k]); H -= p[k] * lp[k]; }
    const float logp = a_s == 0 ? lp[0] : (a_s == 1 ? lp[1] : lp[2]);
    float dlogp;
    float ratio_ppo = 0.f;
    if (ALGO == ALGO_PPO) {
        const float adv_n = hp.normalize_adv ? (adv_s - adv_mean) / adv_den : adv_s;
        const float ratio = GS_EXPF(logp - lp_old);
        ratio_ppo = ratio;
        const float rc = fminf(fmaxf(ratio, hp.clip_lo), hp.clip_hi);
        const float s1 = adv_n * ratio, s2 = adv_n * rc;
        const bool inrange = (ratio >= hp.clip_lo) && (ratio <= hp.clip_hi);
        dlogp = (inrange || s1 < s2) ? -(s1 * invB) : 0.f;
        pm[PM_SURR] += fminf(s1, s2);
        pm[PM_CLIPF] += inrange ? 0.f : 1.f;
        pm[PM_ADVN] += adv_n; pm[PM_ADVN2] = fmaf(adv_n, adv_n, pm[PM_ADVN2]);
        // clipped value loss
        const float v = A == 2 ? out[2] : out[3];
        const float vd = v - v_old;
        const float eu = v - ret_s, lu = eu * eu;
        const float vc = v_old + fminf(fmaxf(vd, -hp.clip_vf), hp.clip_vf);
        const float ec = vc - ret_s, lc = ec * ec;
        const bool in_vf = (vd >= -hp.clip_vf) && (vd <= hp.clip_vf);
        const float gc = in_vf ? 2.f * ec : 0.f;
        const float dv = lu > lc ? 2.f * eu : (lu < lc ? gc : 0.5f * (2.f * eu) + 0.5f * gc);
        const float gv = hp.vf_coef * dv * invB;
        if (A == 2) g[2] = gv; else g[3] = gv;
        pm[PM_VLOSS] += fmaxf(lu, lc);
        pm[PM_CLIPF_VF] += in_vf ? 0.f : 1.f;
        const float rv = ret_s - v;
        pm[PM_RV] += rv; pm[PM_RV2] = fmaf(rv, rv, pm[PM_RV2]);
        pm[PM_R] += ret_s; pm[PM_R2] = fmaf(ret_s, ret_s, pm[PM_R2]);
    } else {
        const float ret_n = hp.normalize_ret ? (ret_s - ret_mean) / ret_den : ret_s;
        const float adv_n = hp.normalize_adv ? (adv_s - adv_mean) / adv_den : adv_s;
        const float tgt = hp.policy_targets == 0 ? ret_n : adv_n;
        dlogp = -(tgt * invB);
        pm[PM_SURR] += logp * tgt;
        pm[PM_TGT] += tgt; pm[PM_TGT2] = fmaf(tgt, tgt, pm[PM_TGT2]);
        pm[PM_ADVN] += adv_n; pm[PM_ADVN2] = fmaf(adv_n, adv_n, pm[PM_ADVN2]);
        pm[PM_RETN] += ret_n; pm[PM_RETN2] = fmaf(ret_n, ret_n, pm[PM_RETN2]);
    }
    const float ec_b = hp.ent_coef * invB;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (k < A) g[k] = dlogp * ((k == a_s ? 1.f : 0.f) - p[k]) + ec_b * p[k] * (lp[k] + H);
    pm[PM_ENT] += H;
    pm[PM_KL] += lp_old - logp;
    const float dlp = logp - lp_old;
    const float dcl = fminf(fmaxf(dlp, -20.f), 20.f);             // utils/torch.py:115-118
    const float r2 = (ALGO == ALGO_PPO && dcl == dlp) ? ratio_ppo : GS_EXPF(dcl);   // same argument as the surrogate's ratio: same bits
    pm[PM_AKL] += (r2 - 1.f) - GS_LOGF(r2);
    pm[PM_COUNT] += 1.f;
}

}  // namespace gs
