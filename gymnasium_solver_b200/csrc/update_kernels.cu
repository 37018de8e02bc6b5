// update_kernels.cu — fused minibatch forward + loss + backward for PPO and REINFORCE, the deterministic gradient /
// metric reduction, the global-norm clip and a graph-capturable Adam.
//
// Replaces: RolloutCollector.slice_trajectories gather (utils/rollout_collector.py:657-682),
//           PPOAgent.losses_for_batch (agents/ppo/ppo_agent.py:21-152), REINFORCEAgent.losses_for_batch
//           (agents/reinforce/reinforce_agent.py:11-88), loss.backward(), BaseModel activation hooks
//           (utils/models.py:121-194), compute_grad_norms + clip_gradients (agents/base_agent.py:591-621).
#include <stdlib.h>
#include <string.h>

#include "mlp_tile.cuh"
#include "update_shared.cuh"

namespace gs {

// ---- the fused kernel -----------------------------------------------------------------------------------------------
template <class C, int ALGO, bool TRACK>
__global__ void __launch_bounds__(kThreads, (C::kSmemFloats * 4 <= 110 * 1024) ? 2 : 1)
update_kernel(MlpDev m, BatchDev b, HpDev hp, const double* __restrict__ adv_mom, const double* __restrict__ ret_mom,
              float* __restrict__ grad_partials /* persist: [grid][P] */, float* __restrict__ grads_atomic /* else: [P] */,
              double* __restrict__ metric_partials /* [grid][PM_N] */, uint32_t* __restrict__ dead, int64_t pstride) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    constexpr int S = C::S, TS = C::TS, HL = C::HL;
    const ParamOffsets po = param_offsets(m.D, C::H1, C::H2, m.A, m.has_value);

    load_resident<C>(sm, m);

    float adv_mean = 0.f, adv_den = 1.f, ret_mean = 0.f, ret_den = 1.f;
    if (hp.normalize_adv) norm_consts(adv_mom, adv_mean, adv_den);
    if (hp.normalize_ret) norm_consts(ret_mom, ret_mean, ret_den);
    const float invB = 1.0f / (float)b.n;

    // persistent wgrad accumulators (registers for the whole kernel when C::kPersist)
    float acc_w2[8][4];
    float acc_h[kNH];
    constexpr int G1 = kThreads / C::H1;        // thread groups splitting the 8 padded inputs of layer 1
    constexpr int R1 = kDP / G1;                // inputs per thread
    constexpr int GL = kThreads / HL;           // thread groups splitting the samples for head / bias sums
    float acc_w1[R1];
    float acc_b1 = 0.f, acc_b2 = 0.f, acc_bh = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc_w2[i][c] = 0.f;
#pragma unroll
    for (int r = 0; r < kNH; ++r) acc_h[r] = 0.f;
#pragma unroll
    for (int r = 0; r < R1; ++r) acc_w1[r] = 0.f;

    float pm[PM_N];
#pragma unroll
    for (int i = 0; i < PM_N; ++i) pm[i] = 0.f;
    ActStats st;
    st.sum[0] = st.sum[1] = st.sumsq[0] = st.sumsq[1] = 0.f;
    __shared__ uint32_t s_dead[512];            // H1 + H2 <= 512
    for (int i = tid; i < C::H1 + C::H2; i += kThreads) s_dead[i] = 0u;
    st.dead = s_dead;

    const int64_t n_tiles = (b.n + S - 1) / S;
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t pos0 = tile * S;
        const int valid_rows = (int)min((int64_t)S, b.n - pos0);
        // ---- gather (one thread per sample) -------------------------------------------------------------------
        int a_s = 0;
        float lp_old = 0.f, v_old = 0.f, adv_s = 0.f, ret_s = 0.f;
        if (tid < S) {
            float x[kDP];
#pragma unroll
            for (int d = 0; d < kDP; ++d) x[d] = 0.f;
            if (tid < valid_rows) {
                const int64_t off = sample_offset(b, pos0 + tid);
                const float* o = b.obs + off * b.D;
                if (b.D == 4) {
                    const float4 v4 = __ldg(reinterpret_cast<const float4*>(o));
                    x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[3] = v4.w;
                } else if (b.D == 2) {
                    const float2 v2 = __ldg(reinterpret_cast<const float2*>(o));
                    x[0] = v2.x; x[1] = v2.y;
                } else {
#pragma unroll
                    for (int d = 0; d < kDP; ++d)
                        if (d < b.D) x[d] = __ldg(o + d);
                }
                a_s = __ldg(b.actions + off);
                lp_old = __ldg(b.logp_old + off);
                adv_s = __ldg(b.adv + off);
                ret_s = __ldg(b.ret + off);
                if (ALGO == ALGO_PPO) v_old = __ldg(b.values_old + off);
            }
            float4* xr = reinterpret_cast<float4*>(sm + C::oXS + tid * kLDX);
            xr[0] = make_float4(x[0], x[1], x[2], x[3]);
            xr[1] = make_float4(x[4], x[5], x[6], x[7]);
        }
        __syncthreads();

        // ---- forward ------------------------------------------------------------------------------------------
        forward_backbone<C, TRACK>(sm, m, valid_rows, &st);

        // ---- heads + loss + dLoss/d(head outputs) (one thread per sample) -------------------------------------
        if (tid < S) {
            float g[kNH] = {0.f, 0.f, 0.f, 0.f};
            if (tid < valid_rows) {
                float out[kNH];
                forward_heads<C>(sm, tid, out);
                sample_loss<ALGO>(out, m.A, a_s, lp_old, v_old, adv_s, ret_s, hp, adv_mean, adv_den, ret_mean, ret_den, invB, g, pm);
            }
            *reinterpret_cast<float4*>(sm + C::oG + tid * kNH) = make_float4(g[0], g[1], g[2], g[3]);
        }
        __syncthreads();

        float* hL = sm + (C::H2 > 0 ? C::oA2 : C::oA1);
        const float* G = sm + C::oG;
        // ---- head wgrad: dWh[r][k] += sum_s g[s][r] * hL[s][k] ---------------------------------------------------
        {
            const int k = tid % HL, sg = tid / HL;
            float loc[kNH] = {0.f, 0.f, 0.f, 0.f};
            for (int s = sg; s < S; s += GL) {
                const float h = hL[s * C::LDL + k];
                const float4 g4 = *reinterpret_cast<const float4*>(G + s * kNH);
                loc[0] = fmaf(g4.x, h, loc[0]); loc[1] = fmaf(g4.y, h, loc[1]);
                loc[2] = fmaf(g4.z, h, loc[2]); loc[3] = fmaf(g4.w, h, loc[3]);
            }
            if (C::kPersist) {
#pragma unroll
                for (int r = 0; r < kNH; ++r) acc_h[r] += loc[r];
            } else {
                for (int r = 0; r < m.A; ++r) atomicAdd(grads_atomic + po.wp + (int64_t)r * HL + k, loc[r]);
                if (m.has_value) atomicAdd(grads_atomic + po.wv + k, loc[m.A]);
            }
            if (tid < kNH) {  // head biases
                float sgm = 0.f;
                for (int s = 0; s < S; ++s) sgm += G[s * kNH + tid];
                if (C::kPersist) acc_bh += sgm;
                else if (tid < m.A) atomicAdd(grads_atomic + po.bp + tid, sgm);
                else if (tid == m.A && m.has_value) atomicAdd(grads_atomic + po.bv, sgm);
            }
        }
        __syncthreads();
        // ---- dZL = (G @ Wh) * act'(hL), in place over hL -------------------------------------------------------------
        {
            const float* wh = sm + C::oWH;
#pragma unroll 1
            for (int nc = 0; nc < HL / 64; ++nc) {
                const int col = nc * 64 + 4 * tx;
                float4 w[kNH];
#pragma unroll
                for (int r = 0; r < kNH; ++r) w[r] = *reinterpret_cast<const float4*>(wh + r * C::LDL + col);
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    const int s = ty * TS + j;
                    const float4 g4 = *reinterpret_cast<const float4*>(G + s * kNH);
                    float4* hp4 = reinterpret_cast<float4*>(hL + s * C::LDL + col);
                    const float4 h = *hp4;
                    float4 d;
                    d.x = g4.x * w[0].x + g4.y * w[1].x + g4.z * w[2].x + g4.w * w[3].x;
                    d.y = g4.x * w[0].y + g4.y * w[1].y + g4.z * w[2].y + g4.w * w[3].y;
                    d.z = g4.x * w[0].z + g4.y * w[1].z + g4.z * w[2].z + g4.w * w[3].z;
                    d.w = g4.x * w[0].w + g4.y * w[1].w + g4.z * w[2].w + g4.w * w[3].w;
                    d.x *= act_bwd(h.x, m.act); d.y *= act_bwd(h.y, m.act);
                    d.z *= act_bwd(h.z, m.act); d.w *= act_bwd(h.w, m.act);
                    *hp4 = d;
                }
            }
        }
        __syncthreads();

        if (C::H2 > 0) {
            const float* dZ2 = sm + C::oA2;
            float* a1 = sm + C::oA1;
            // ---- layer-2 bias + weight gradients -----------------------------------------------------------------------
            {
                const int n = tid % C::HL, sg = tid / C::HL;
                float sgm = 0.f;
                for (int s = sg; s < S; s += GL) sgm += dZ2[s * C::LD2 + n];
                if (C::kPersist) acc_b2 += sgm;
                else atomicAdd(grads_atomic + po.b2 + n, sgm);
            }
            {
                const int half = tid >> 7, t = tid & 127, tn = t >> 4, tk = t & 15;
                if (C::kPersist) {
                    tile_tn<S>(acc_w2, dZ2, C::LD2, a1, C::LD1, tn, tk, half);
                } else {
#pragma unroll 1
                    for (int nb = 0; nb < C::H2 / 64; ++nb)
#pragma unroll 1
                        for (int kb = 0; kb < C::H1 / 64; ++kb) {
                            float acc[8][4];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
#pragma unroll
                                for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
                            tile_tn<S>(acc, dZ2 + nb * 64, C::LD2, a1 + kb * 64, C::LD1, tn, tk, half);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    atomicAdd(grads_atomic + po.w2 + (int64_t)(nb * 64 + 8 * tn + i) * C::H1 + kb * 64 + 4 * tk + c, acc[i][c]);
                        }
                }
            }
            __syncthreads();
            // ---- dZ1 = (dZ2 @ W2) * act'(h1), in place over a1 ------------------------------------------------------------
#pragma unroll 1
            for (int kc = 0; kc < C::H1 / 64; ++kc) {
                float acc[TS][4];
#pragma unroll
                for (int j = 0; j < TS; ++j)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
#pragma unroll 1
                for (int nc = 0; nc < C::H2 / 64; ++nc) {
                    if (!C::kResidentW2) {
                        __syncthreads();
                        stage_w64(sm + C::oWB, m.w2, C::H1, nc * 64, kc * 64);
                        __syncthreads();
                    }
                    tile_nn<TS>(acc, dZ2 + nc * 64, C::LD2, sm + C::oWB, kWLD, tx, ty);
                }
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    float4* hp4 = reinterpret_cast<float4*>(a1 + (ty * TS + j) * C::LD1 + kc * 64 + 4 * tx);
                    const float4 h = *hp4;
                    *hp4 = make_float4(acc[j][0] * act_bwd(h.x, m.act), acc[j][1] * act_bwd(h.y, m.act),
                                       acc[j][2] * act_bwd(h.z, m.act), acc[j][3] * act_bwd(h.w, m.act));
                }
            }
            __syncthreads();
        }
        // ---- layer-1 weight / bias gradients: dW1[n][d] += sum_s dZ1[s][n] * x[s][d] --------------------------------------
        {
            const float* dZ1 = sm + C::oA1;
            const float* xs = sm + C::oXS;
            const int n = tid % C::H1, dg = tid / C::H1;
            float loc[R1];
#pragma unroll
            for (int r = 0; r < R1; ++r) loc[r] = 0.f;
            float sb = 0.f;
            for (int s = 0; s < S; ++s) {
                const float dz = dZ1[s * C::LD1 + n];
#pragma unroll
                for (int r = 0; r < R1; ++r) loc[r] = fmaf(dz, xs[s * kLDX + dg * R1 + r], loc[r]);
                sb += dz;
            }
            if (C::kPersist) {
#pragma unroll
                for (int r = 0; r < R1; ++r) acc_w1[r] += loc[r];
                if (dg == 0) acc_b1 += sb;
            } else {
#pragma unroll
                for (int r = 0; r < R1; ++r)
                    if (dg * R1 + r < m.D) atomicAdd(grads_atomic + po.w1 + (int64_t)n * m.D + dg * R1 + r, loc[r]);
                if (dg == 0) atomicAdd(grads_atomic + po.b1 + n, sb);
            }
        }
        __syncthreads();
    }

    // ---- fold the register accumulators of this CTA into one partial gradient vector (deterministic order) ----------
    if (C::kPersist) {
        float* gsm = sm;  // reuse the activation tiles: P floats
        const int P = (int)po.total;
        for (int i = tid; i < P; i += kThreads) gsm[i] = 0.f;
        __syncthreads();
        const int half = tid >> 7, t = tid & 127, tn = t >> 4, tk = t & 15;
        const int kh = tid % HL, sg = tid / HL;
        constexpr int ROUNDS = GL > 2 ? GL : 2;
        for (int round = 0; round < ROUNDS; ++round) {
            if (C::H2 > 0 && half == round) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) gsm[po.w2 + (8 * tn + i) * C::H1 + 4 * tk + c] += acc_w2[i][c];
            }
            if (sg == round) {
                for (int r = 0; r < m.A; ++r) gsm[po.wp + r * HL + kh] += acc_h[r];
                if (m.has_value) gsm[po.wv + kh] += acc_h[m.A];
                if (C::H2 > 0) gsm[po.b2 + kh] += acc_b2;
            }
            __syncthreads();
        }
        {
            const int n = tid % C::H1, dg = tid / C::H1;
#pragma unroll
            for (int r = 0; r < R1; ++r)
                if (dg * R1 + r < m.D) gsm[po.w1 + n * m.D + dg * R1 + r] = acc_w1[r];
            if (dg == 0) gsm[po.b1 + n] = acc_b1;
            if (tid < m.A) gsm[po.bp + tid] = acc_bh;
            else if (tid == m.A && m.has_value) gsm[po.bv] = acc_bh;
        }
        __syncthreads();
        float* out = grad_partials + (size_t)blockIdx.x * pstride;
        for (int i = tid; i < P; i += kThreads) out[i] = gsm[i];
    }

    // ---- dead-unit counts of this CTA: one global integer atomic per unit --------------------------------------------------
    if (TRACK) {
        __syncthreads();
        for (int i = tid; i < C::H1 + C::H2; i += kThreads)
            if (s_dead[i]) atomicAdd(dead + i, s_dead[i]);
    }
    // ---- metric partials: warp shuffle, then one cross-warp pass through shared memory ---------------------------------
    if (TRACK) { pm[PM_Z0] = st.sum[0]; pm[PM_Z0SQ] = st.sumsq[0]; pm[PM_Z1] = st.sum[1]; pm[PM_Z1SQ] = st.sumsq[1]; }
    __syncthreads();
    {
        double* red = reinterpret_cast<double*>(sm + ((C::kPersist ? (int)po.total : 0) + 3) / 4 * 4 + 4);  // past gsm, 16 B aligned
        const int lane = tid & 31, w = tid >> 5;
#pragma unroll
        for (int i = 0; i < PM_N; ++i) {
            const double v = warp_sum((double)pm[i]);
            if (lane == 0) red[w * PM_N + i] = v;
        }
        __syncthreads();
        if (tid < PM_N) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < kThreads / 32; ++ww) s += red[ww * PM_N + tid];
            metric_partials[(size_t)blockIdx.x * PM_N + tid] = s;
        }
    }
}

// ---- stage 2: ordered reduction over CTAs + metric finalisation ---------------------------------------------------------
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_cta, int64_t P, int64_t pstride, float* __restrict__ grads) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += partials[(size_t)c * pstride + i];
    grads[i] = s;
}

__device__ __forceinline__ void finalize_metrics_body(const double* __restrict__ partials, int n_cta, int algo, int H1, int H2, int track,
                                                      float vf_coef, float ent_coef, int normalize_adv, int normalize_ret,
                                                      const uint32_t* __restrict__ dead, double* __restrict__ metrics) {
    // needs blockDim.x >= 256.  Fixed summation order (8 strided groups per metric, then the groups in order): deterministic.
    constexpr int kGroups = 8;
    __shared__ double grp[PM_N * kGroups];
    __shared__ double tot[PM_N];
    __shared__ double dead_stats[4];
    const int tid = threadIdx.x;
    if (tid < PM_N * kGroups) {
        const int q = tid % PM_N, g = tid / PM_N;
        double s = 0.0;
#pragma unroll 4
        for (int c = g; c < n_cta; c += kGroups) s += __ldcg(partials + (size_t)c * PM_N + q);
        grp[g * PM_N + q] = s;
    }
    if (track && tid >= 192 && tid < 256) {   // dead-neuron fractions per hooked layer: mean and max over neurons (integer counts: order-free)
        const int l = (tid - 192) >> 5, lane = tid & 31;
        const int H = l == 0 ? H1 : H2, base = l == 0 ? 0 : H1;
        unsigned long long sum = 0; uint32_t mx = 0;
        for (int n = lane; n < H; n += 32) { const uint32_t c = __ldcg(dead + base + n); sum += c; mx = c > mx ? c : mx; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const uint32_t m2 = __shfl_xor_sync(0xffffffffu, mx, o);
            mx = m2 > mx ? m2 : mx;
        }
        if (lane == 0) { dead_stats[2 * l] = (double)sum; dead_stats[2 * l + 1] = (double)mx; }
    }
    __syncthreads();
    if (tid < PM_N) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) s += grp[g * PM_N + tid];
        tot[tid] = s;
    }
    __syncthreads();
    if (tid != 0) return;       // leaves this (inlined) function only: a caller with more work resumes with every thread
    const double B = tot[PM_COUNT];
    const double ent = tot[PM_ENT] / B;
    auto ustd = [](double s, double s2, double n) { double v = (s2 - s * s / n) / (n - 1.0); return sqrt(v > 0.0 ? v : 0.0); };
    for (int i = 0; i < GS_M_GRAD_NORM_ALL; ++i) metrics[i] = 0.0;
    metrics[GS_M_RET_NORM_MEAN] = metrics[GS_M_RET_NORM_STD] = 0.0;
    metrics[GS_M_BATCH_COUNT] = B;
    const double policy_loss = -tot[PM_SURR] / B;
    metrics[GS_M_LOSS_POLICY] = policy_loss;
    metrics[GS_M_LOSS_ENTROPY] = -ent;
    metrics[GS_M_ENTROPY] = ent;
    metrics[GS_M_LOSS_ENTROPY_SCALED] = (double)ent_coef * -ent;
    metrics[GS_M_KL] = tot[PM_KL] / B;
    metrics[GS_M_APPROX_KL] = tot[PM_AKL] / B;
    if (algo == ALGO_PPO) {
        const double vl = tot[PM_VLOSS] / B;
        metrics[GS_M_LOSS_VALUE] = vl;
        metrics[GS_M_LOSS_VALUE_SCALED] = (double)vf_coef * vl;
        metrics[GS_M_LOSS_TOTAL] = policy_loss + (double)vf_coef * vl + (double)ent_coef * -ent;
        metrics[GS_M_CLIP_FRACTION] = tot[PM_CLIPF] / B;
        metrics[GS_M_CLIP_FRACTION_VF] = tot[PM_CLIPF_VF] / B;
        const double var_rv = (tot[PM_RV2] - tot[PM_RV] * tot[PM_RV] / B) / (B - 1.0);
        const double var_r = (tot[PM_R2] - tot[PM_R] * tot[PM_R] / B) / (B - 1.0);
        metrics[GS_M_EXPLAINED_VAR] = 1.0 - var_rv / var_r;
    } else {
        metrics[GS_M_LOSS_TOTAL] = policy_loss + (double)ent_coef * -ent;
        metrics[GS_M_TARGETS_MEAN] = tot[PM_TGT] / B;
        metrics[GS_M_TARGETS_STD] = ustd(tot[PM_TGT], tot[PM_TGT2], B);
        if (normalize_ret) {
            metrics[GS_M_RET_NORM_MEAN] = tot[PM_RETN] / B;
            metrics[GS_M_RET_NORM_STD] = ustd(tot[PM_RETN], tot[PM_RETN2], B);
        }
    }
    if (normalize_adv) {
        metrics[GS_M_ADV_NORM_MEAN] = tot[PM_ADVN] / B;
        metrics[GS_M_ADV_NORM_STD] = ustd(tot[PM_ADVN], tot[PM_ADVN2], B);
    }
    if (track) {
        const double n0 = B * H1;
        metrics[GS_M_ACT0_MEAN] = tot[PM_Z0] / n0;
        metrics[GS_M_ACT0_STD] = ustd(tot[PM_Z0], tot[PM_Z0SQ], n0);
        metrics[GS_M_ACT0_DEAD_PCT] = dead_stats[0] / n0;
        metrics[GS_M_ACT0_DEAD_MAX] = dead_stats[1] / B;
        if (H2 > 0) {
            const double n1 = B * H2;
            metrics[GS_M_ACT1_MEAN] = tot[PM_Z1] / n1;
            metrics[GS_M_ACT1_STD] = ustd(tot[PM_Z1], tot[PM_Z1SQ], n1);
            metrics[GS_M_ACT1_DEAD_PCT] = dead_stats[2] / n1;
            metrics[GS_M_ACT1_DEAD_MAX] = dead_stats[3] / B;
        }
    }
}

// One launch folds the per-CTA partials: blocks 0 .. n_red-1 sum the partial gradient vectors in CTA order (deterministic),
// the last block turns the metric partials into the metric vector.
__global__ void reduce_and_finalize_kernel(const float* __restrict__ partials, int n_cta, int64_t P, int64_t pstride, float* __restrict__ grads,
                                           const double* __restrict__ metric_partials, int n_metric_cta, int algo, int H1, int H2, int track, float vf_coef,
                                           float ent_coef, int normalize_adv, int normalize_ret, const uint32_t* __restrict__ dead,
                                           double* __restrict__ metrics) {
    if (blockIdx.x == gridDim.x - 1) {
        finalize_metrics_body(metric_partials, n_metric_cta, algo, H1, H2, track, vf_coef, ent_coef, normalize_adv, normalize_ret, dead, metrics);
        return;
    }
    if (!partials) return;                      // atomic-accumulation configurations have no partial vectors
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += partials[(size_t)c * pstride + i];
    grads[i] = s;
}

// ---- fused step tail (gs_update_finish): partial reduction -> NVLink gradient exchange -> metrics -> clip -> Adam ----------
struct PeerDev {
    int rank, world;
    unsigned long long* slots[GS_PEER_MAX_WORLD];   // slots[r]: rank r's receive buffer [2 phases][world][stride] of {value bits, epoch} words (peer-mapped)
    uint32_t epoch;                       // call counter (>= 1), identical on every rank
    int64_t stride;
};
struct AdamDev { float *p, *m, *v; int64_t* step; float lr, beta1, beta2, eps; };
struct FinishDev {
    const float* partials; int n_cta; int64_t P, pstride;
    const double* metric_partials; int n_metric_cta;
    int algo, H1, H2, track, normalize_adv, normalize_ret; float vf_coef, ent_coef, max_norm;
    const uint32_t* dead; uint32_t* ticket;
    int split;                  // large models: this kernel only reduces + posts; finish_receive_kernel / finish_apply_kernel do the rest grid-wide
};

// The exchange carries its own arrival flags: every gradient element travels as ONE 8-byte word {fp32 bits, call epoch} (an aligned 8-byte store
// is not torn), so the receiver polls the data itself -- no system-wide fence after the stores, no separate flag store behind a release, no
// second NVLink round trip (the protocol NCCL calls LL).  Measured on 2 GPUs with per-phase globaltimer stamps: fenced stores + flag exchange
// cost 12 + 8 us per minibatch in the step tail against 6 us for the same phase on one GPU.
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

#ifdef GS_FINISH_TRACE   // development aid: globaltimer (ns) sums over calls of the last block's phases; gs_debug_finish_trace() reads them
__device__ unsigned long long g_fin_trace[8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GS_FT(k) do { if (tid == 0) { const unsigned long long now_ = gtime(); atomicAdd(&g_fin_trace[k], now_ - t_prev_); t_prev_ = now_; } } while (0)
#else
#define GS_FT(k) do { } while (0)
#endif
constexpr int64_t kFinishSplitP = 8192;  // models above this finish grid-wide (finish_receive_kernel / finish_apply_kernel)
constexpr int kFinishThreads = 1024;    // phase A: 64 parameters x 16 groups of partial vectors per block
constexpr int kFinishGroups = kFinishThreads / 64;
__global__ void __launch_bounds__(kFinishThreads) update_finish_kernel(FinishDev f, ParamOffsets po, float* __restrict__ grads, AdamDev ad, PeerDev peer,
                                                                       double* __restrict__ metrics, double* __restrict__ metrics_sum) {
    __shared__ double scratch[32];
    __shared__ double sq[3];
    __shared__ float adam_c[2];
    __shared__ int is_last;
    const int tid = threadIdx.x;
    const int phase = (int)(peer.epoch & 1u);
#ifdef GS_FINISH_TRACE
    unsigned long long t_prev_ = gtime();
#endif
    // programmatic dependent launch: the next minibatch's update kernel may be scheduled now; it runs its weight-independent
    // prologue and then blocks in griddepcontrol.wait until this grid has completed and flushed (parameters, moments, ticket)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 64 && ad.p) {                                   // bias corrections of THIS step, from the step count before it: two fp64 pow()
        const int64_t step = *ad.step + 1;                      // calls that would otherwise sit in the serial phase of the last block
        const double bc1 = 1.0 - pow((double)ad.beta1, (double)step);
        const double bc2 = 1.0 - pow((double)ad.beta2, (double)step);
        adam_c[0] = (float)((double)ad.lr / bc1);
        adam_c[1] = (float)sqrt(bc2);
    }
    // The metric vector of the step (slots below GS_M_GRAD_NORM_ALL, batch count, return-normalisation slots) depends only on the update
    // kernel's metric partials: one EXTRA block (the grid's last; it takes no ticket) computes it beside the gradient path instead of the
    // last reducing block after it (5 of that block's 12 serial microseconds).  Large models: finish_apply_kernel's block 0 does it.
    const unsigned n_reduce = f.split ? gridDim.x : gridDim.x - 1;
    if (blockIdx.x >= n_reduce) {
        finalize_metrics_body(f.metric_partials, f.n_metric_cta, f.algo, f.H1, f.H2, f.track, f.vf_coef, f.ent_coef, f.normalize_adv, f.normalize_ret,
                              f.dead, metrics);
        __syncthreads();
        if (metrics_sum && tid < GS_M_SCRATCH && (tid < GS_M_GRAD_NORM_ALL || tid > GS_M_CLIP_COEF)) metrics_sum[tid] += metrics[tid];
        if (f.track)                                            // consumed: leave the dead-unit counters zeroed for the next step
            for (int i = tid; i < f.H1 + f.H2; i += kFinishThreads) const_cast<uint32_t*>(f.dead)[i] = 0u;
        return;
    }
    // ---- phase A (every block): sum of the per-CTA partial vectors in a fixed order (16 contiguous groups of CTAs, then the
    //      groups in order); the result goes to grads (one rank) or straight into every rank's receive slot over NVLink
    //      (posted P2P stores, no reads cross the link)
    {
        __shared__ float partg[kFinishGroups][64];
        const int pl = tid & 63, cg = tid >> 6;
        const int64_t i = (int64_t)blockIdx.x * 64 + pl;
        float s = 0.f;
        if (i < f.P && f.partials) {
            const int per = (f.n_cta + kFinishGroups - 1) / kFinishGroups;
            const int c0 = cg * per, c1 = min(f.n_cta, c0 + per);
#pragma unroll 8
            for (int c = c0; c < c1; ++c) s += __ldcg(f.partials + (size_t)c * f.pstride + i);
        }
        partg[cg][pl] = s;
        __syncthreads();
        if (cg == 0 && i < f.P) {
            if (f.partials) {
#pragma unroll
                for (int g = 1; g < kFinishGroups; ++g) s += partg[g][pl];
            } else {
                s = grads[i];                                   // atomic-accumulation configurations
            }
            if (peer.world > 1) {
                const int64_t o = ((int64_t)phase * peer.world + peer.rank) * peer.stride + i;
                const unsigned long long word = ((unsigned long long)peer.epoch << 32) | (unsigned long long)__float_as_uint(s);
                for (int r = 0; r < peer.world; ++r) st_relaxed_sys_u64(peer.slots[r] + o, word);
            } else {
                grads[i] = s;
            }
        }
    }
    if (peer.world == 1) __threadfence();                       // several ranks: the last block polls the gradient words themselves
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(f.ticket, 1u) == n_reduce - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    if (f.split) { if (tid == 0) *f.ticket = 0u; return; }
    __threadfence();
    GS_FT(0);                                                  // entry of the last block -> every block's phase A done
    // ---- phase B (the last block to arrive) -----------------------------------------------------------------------------------
    if (peer.world > 1) {
        // every rank's word of element i, in rank order (every rank computes the bit-identical mean); a word is valid once it carries this
        // call's epoch.  Bounded spin -> trap: a rank left the lock-step call sequence (fail, do not hang).
        const unsigned long long* mine = peer.slots[peer.rank] + (int64_t)phase * peer.world * peer.stride;
        const float inv_w = 1.0f / (float)peer.world;
        for (int64_t i = tid; i < f.P; i += kFinishThreads) {
            unsigned long long w[GS_PEER_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < GS_PEER_MAX_WORLD; ++r)          // all ranks' words requested at once
                if (r < peer.world) w[r] = ld_relaxed_sys_u64(mine + (int64_t)r * peer.stride + i);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < GS_PEER_MAX_WORLD; ++r) {
                if (r < peer.world) {
                    uint32_t spin = 0;
                    while ((uint32_t)(w[r] >> 32) != peer.epoch) {
                        if (++spin > (1u << 27)) __trap();
                        __nanosleep(20);
                        w[r] = ld_relaxed_sys_u64(mine + (int64_t)r * peer.stride + i);
                    }
                    s += __uint_as_float((uint32_t)w[r]);
                }
            }
            grads[i] = s * inv_w;
        }
        __syncthreads();
        GS_FT(2);                                              // every rank's words received and summed
    }
    // group norms in a fixed order (utils/models.py:196-230), clip coefficient (torch.nn.utils.clip_grad_norm_).  This kernel serves models of
    // at most kFinishSplitP parameters: a thread owns <= 8 elements, requested together (one L2 round trip, not one per element) and kept
    // in registers for the optimizer step.
    constexpr int kPer = (int)(kFinishSplitP / kFinishThreads);
    float g[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int64_t i = tid + (int64_t)k * kFinishThreads;
        g[k] = i < po.total ? __ldcg(grads + i) : 0.f;          // other blocks wrote it in phase A: read through L2
    }
    double part[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int64_t i = tid + (int64_t)k * kFinishThreads;
        const double v = (double)g[k];
        const int grp = i < po.wp ? 0 : (i < po.wv ? 1 : 2);
        part[grp] += v * v;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double t = block_sum(part[k], scratch);
        if (tid == 0) sq[k] = t;
    }
    if (tid == 64 && ad.p) *ad.step = *ad.step + 1;           // adam_c was computed from the old count at kernel entry
    __syncthreads();
    const double total = sqrt(sq[0] + sq[1] + sq[2]);
    double coef = 1.0;
    if (f.max_norm > 0.f) { coef = (double)f.max_norm / (total + 1e-6); coef = coef > 1.0 ? 1.0 : coef; }
    if (tid == 0) {
        metrics[GS_M_GRAD_NORM_ALL] = total;
        metrics[GS_M_GRAD_NORM_BACKBONE] = sqrt(sq[0]);
        metrics[GS_M_GRAD_NORM_POLICY] = sqrt(sq[1]);
        metrics[GS_M_GRAD_NORM_VALUE] = sqrt(sq[2]);
        metrics[GS_M_CLIP_COEF] = coef;
        *f.ticket = 0u;
    }
    const float c = (float)coef;
    const bool scale = coef < 1.0;
    if (ad.p) {
        const float step_size = adam_c[0], bc2_sqrt = adam_c[1];
        float* __restrict__ pp = ad.p;
        float* __restrict__ pm = ad.m;
        float* __restrict__ pv = ad.v;
        float m0[kPer], v0[kPer], p0[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int64_t i = tid + (int64_t)k * kFinishThreads;
            if (i < po.total) { m0[k] = pm[i]; v0[k] = pv[i]; p0[k] = pp[i]; }
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int64_t i = tid + (int64_t)k * kFinishThreads;
            if (i < po.total) {
                float gi = g[k];
                if (scale) { gi *= c; grads[i] = gi; }
                const float mi = m0[k] + (gi - m0[k]) * (1.f - ad.beta1);
                const float vi = v0[k] * ad.beta2 + (1.f - ad.beta2) * gi * gi;
                pm[i] = mi; pv[i] = vi;
                const float denom = sqrtf(vi) / bc2_sqrt + ad.eps;
                pp[i] = p0[k] - step_size * (mi / denom);
            }
        }
    } else if (scale) {
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int64_t i = tid + (int64_t)k * kFinishThreads;
            if (i < po.total) grads[i] = g[k] * c;
        }
    }
    __syncthreads();
    GS_FT(3);                                                  // metrics, norms, clip, Adam
#ifdef GS_FINISH_TRACE
    if (tid == 0) atomicAdd(&g_fin_trace[7], 1ull);
#endif
    if (metrics_sum && tid >= GS_M_GRAD_NORM_ALL && tid <= GS_M_CLIP_COEF) metrics_sum[tid] += metrics[tid];   // the slots this block wrote
}

// ---- the step tail of LARGE models, grid-wide (P > kFinishSplitP) ---------------------------------------------------------------------
// One block finishing a 68 k-parameter model (256 x 256) took 137 us per minibatch -- 7 % of a MountainCar-v0 iteration, and more than the
// sharded update kernel itself at 8 GPUs.  update_finish_kernel then only reduces the partial vectors (and posts the words to the peers);
//   finish_receive_kernel: every block receives / reads its 1024 gradient elements and writes the squared-norm partials of the three groups;
//   finish_apply_kernel:   every block sums those partials in block order (identical everywhere), clips and applies Adam to its elements;
//                          block 0 also finalises the metric vector.
__global__ void __launch_bounds__(kFinishThreads) finish_receive_kernel(FinishDev f, ParamOffsets po, float* __restrict__ grads, AdamDev ad, PeerDev peer,
                                                                        double* __restrict__ sq_part /* [gridDim.x][3] + step snapshot */) {
    __shared__ double scratch[32];
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * kFinishThreads + tid;
    float g = 0.f;
    if (i < f.P) {
        if (peer.world > 1) {
            const int phase = (int)(peer.epoch & 1u);
            const unsigned long long* mine = peer.slots[peer.rank] + (int64_t)phase * peer.world * peer.stride;
            unsigned long long w[GS_PEER_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < GS_PEER_MAX_WORLD; ++r)
                if (r < peer.world) w[r] = ld_relaxed_sys_u64(mine + (int64_t)r * peer.stride + i);
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < GS_PEER_MAX_WORLD; ++r) {
                if (r < peer.world) {
                    uint32_t spin = 0;
                    while ((uint32_t)(w[r] >> 32) != peer.epoch) {
                        if (++spin > (1u << 27)) __trap();
                        __nanosleep(20);
                        w[r] = ld_relaxed_sys_u64(mine + (int64_t)r * peer.stride + i);
                    }
                    s += __uint_as_float((uint32_t)w[r]);   // rank order: identical on every rank
                }
            }
            g = s * (1.0f / (float)peer.world);
            grads[i] = g;
        } else {
            g = __ldcg(grads + i);
        }
    }
    const int grp = i < po.wp ? 0 : (i < po.wv ? 1 : 2);
    const double v = (double)g * (double)g;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double t = block_sum(grp == k ? v : 0.0, scratch);
        if (tid == 0) sq_part[(int64_t)blockIdx.x * 3 + k] = t;
    }
    if (blockIdx.x == 0 && tid == 0 && ad.p) reinterpret_cast<int64_t*>(sq_part + (int64_t)gridDim.x * 3)[0] = *ad.step + 1;   // this step's count
}

__global__ void __launch_bounds__(kFinishThreads) finish_apply_kernel(FinishDev f, ParamOffsets po, float* __restrict__ grads, AdamDev ad,
                                                                      const double* __restrict__ sq_part, double* __restrict__ metrics,
                                                                      double* __restrict__ metrics_sum) {
    __shared__ double sq[3];
    __shared__ float adam_c[2];
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next minibatch's update kernel may start its prologue
    if (tid < 3) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += sq_part[(int64_t)b * 3 + tid];   // block order: the same sum in every block
        sq[tid] = t;
    }
    const int64_t step = ad.p ? reinterpret_cast<const int64_t*>(sq_part + (int64_t)gridDim.x * 3)[0] : 0;
    if (tid == 64 && ad.p) {
        const double bc1 = 1.0 - pow((double)ad.beta1, (double)step);
        const double bc2 = 1.0 - pow((double)ad.beta2, (double)step);
        adam_c[0] = (float)((double)ad.lr / bc1);
        adam_c[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const double total = sqrt(sq[0] + sq[1] + sq[2]);
    double coef = 1.0;
    if (f.max_norm > 0.f) { coef = (double)f.max_norm / (total + 1e-6); coef = coef > 1.0 ? 1.0 : coef; }
    const float c = (float)coef;
    const bool scale = coef < 1.0;
    const int64_t i = (int64_t)blockIdx.x * kFinishThreads + tid;
    if (i < po.total) {
        float gi = __ldcg(grads + i);
        if (scale) { gi *= c; grads[i] = gi; }
        if (ad.p) {
            const float m0 = ad.m[i], v0 = ad.v[i], p0 = ad.p[i];
            const float mi = m0 + (gi - m0) * (1.f - ad.beta1);
            const float vi = v0 * ad.beta2 + (1.f - ad.beta2) * gi * gi;
            ad.m[i] = mi; ad.v[i] = vi;
            const float denom = sqrtf(vi) / adam_c[1] + ad.eps;
            ad.p[i] = p0 - adam_c[0] * (mi / denom);
        }
    }
    if (blockIdx.x != 0) return;
    if (tid == 64 && ad.p) *ad.step = step;
    finalize_metrics_body(f.metric_partials, f.n_metric_cta, f.algo, f.H1, f.H2, f.track, f.vf_coef, f.ent_coef, f.normalize_adv, f.normalize_ret,
                          f.dead, metrics);
    if (tid == 0) {
        metrics[GS_M_GRAD_NORM_ALL] = total;
        metrics[GS_M_GRAD_NORM_BACKBONE] = sqrt(sq[0]);
        metrics[GS_M_GRAD_NORM_POLICY] = sqrt(sq[1]);
        metrics[GS_M_GRAD_NORM_VALUE] = sqrt(sq[2]);
        metrics[GS_M_CLIP_COEF] = coef;
    }
    __syncthreads();
    if (metrics_sum && tid < GS_M_SCRATCH) metrics_sum[tid] += metrics[tid];
    if (f.track)
        for (int i = tid; i < f.H1 + f.H2; i += kFinishThreads) const_cast<uint32_t*>(f.dead)[i] = 0u;
}

// ---- sum over ranks of a short fp64 vector through the peer buffers (gs_peer_allreduce_f64) --------------------------------------
// The minibatch moments of a pass (6 doubles per minibatch) have to be those of the GLOBAL minibatch.  An NCCL all-reduce of that vector
// costs ~150 us on a side stream AND its kernel cannot share an SM with a persistent update CTA (registers), so it ran between two update
// kernels and delayed one CTA -- hence the whole next launch -- by its duration: ~170 us per pass on 4 GPUs.  This one is a single block of
// 256 threads and 30 registers that co-resides with the update kernel: every double travels as two {32-bit half, epoch} words (the protocol
// of the gradient exchange), summed in rank order so every rank holds the identical result.
constexpr int64_t kPeerMomDoubles = 8192;
struct PeerMomDev {
    int rank, world;
    unsigned long long* slots[GS_PEER_MAX_WORLD];   // [2 phases][world][2 * kPeerMomDoubles] words
    uint32_t epoch;
};
__device__ __forceinline__ uint32_t poll_word(const unsigned long long* src, uint32_t epoch) {
    unsigned long long w = ld_relaxed_sys_u64(src);
    uint32_t spin = 0;
    while ((uint32_t)(w >> 32) != epoch) {
        if (++spin > (1u << 27)) __trap();                      // a rank left the lock-step call sequence
        __nanosleep(20);
        w = ld_relaxed_sys_u64(src);
    }
    return (uint32_t)w;
}
__global__ void __launch_bounds__(256) peer_allreduce_f64_kernel(PeerMomDev pm, double* __restrict__ data, int n) {
    const int tid = threadIdx.x;
    const int64_t stride = 2 * kPeerMomDoubles;
    const int64_t base = (int64_t)(pm.epoch & 1u) * pm.world * stride;
    const uint32_t* halves = reinterpret_cast<const uint32_t*>(data);
    for (int w = tid; w < 2 * n; w += 256) {
        const unsigned long long word = ((unsigned long long)pm.epoch << 32) | (unsigned long long)halves[w];
        for (int r = 0; r < pm.world; ++r) st_relaxed_sys_u64(pm.slots[r] + base + (int64_t)pm.rank * stride + w, word);
    }
    // own words are polled like everybody else's, so data[i] is overwritten only after both of its halves were read and posted
    const unsigned long long* mine = pm.slots[pm.rank] + base;
    for (int i = tid; i < n; i += 256) {
        double s = 0.0;
        for (int r = 0; r < pm.world; ++r) {
            const uint32_t lo = poll_word(mine + (int64_t)r * stride + 2 * i, pm.epoch);
            const uint32_t hi = poll_word(mine + (int64_t)r * stride + 2 * i + 1, pm.epoch);
            s += __hiloint2double((int)hi, (int)lo);
        }
        data[i] = s;
    }
}

// ---- minibatch moments of a rollout field (advantage / return batch normalisation) -------------------------------------
__global__ void batch_moments_kernel(BatchDev b, const float* __restrict__ field, double* __restrict__ out) {
    __shared__ double scratch[32];
    double s = 0.0, s2 = 0.0;
    for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < b.n; pos += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)__ldg(field + sample_offset(b, pos));
        s += v; s2 += v * v;
    }
    s = block_sum(s, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(out + 0, s);
        atomicAdd(out + 1, s2);
        if (blockIdx.x == 0) atomicAdd(out + 2, (double)b.n);
    }
}

// ---- grad norms (all / backbone / policy_head / value_head) + global-norm clip -----------------------------------------
__global__ void grad_norm_kernel(const float* __restrict__ g, ParamOffsets po, double* __restrict__ sq /* [3] zeroed */) {
    __shared__ double scratch[32];
    double part[3] = {0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < po.total; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)g[i];
        const int grp = i < po.wp ? 0 : (i < po.wv ? 1 : 2);
        part[grp] += v * v;
    }
    for (int k = 0; k < 3; ++k) {
        const double v = block_sum(part[k], scratch);
        if (threadIdx.x == 0 && v != 0.0) atomicAdd(sq + k, v);
    }
}

__global__ void clip_scale_kernel(float* __restrict__ g, int64_t P, const double* __restrict__ sq, float max_norm,
                                  double* __restrict__ metrics) {
    const double total = sqrt(sq[0] + sq[1] + sq[2]);
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total + 1e-6), clamped to 1
    double coef = 1.0;
    if (max_norm > 0.f) { coef = (double)max_norm / (total + 1e-6); coef = coef > 1.0 ? 1.0 : coef; }
    if (blockIdx.x == 0 && threadIdx.x == 0 && metrics) {
        metrics[GS_M_GRAD_NORM_ALL] = total;
        metrics[GS_M_GRAD_NORM_BACKBONE] = sqrt(sq[0]);
        metrics[GS_M_GRAD_NORM_POLICY] = sqrt(sq[1]);
        metrics[GS_M_GRAD_NORM_VALUE] = sqrt(sq[2]);
        metrics[GS_M_CLIP_COEF] = coef;
    }
    if (coef < 1.0) {
        const float c = (float)coef;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) g[i] *= c;
    }
}

// grad norms + clip of a small parameter vector in ONE block: group sums of squares in a fixed order, then the scaling pass
__global__ void __launch_bounds__(1024) clip_grad_norm_kernel(float* __restrict__ g, ParamOffsets po, float max_norm, double* __restrict__ metrics) {
    __shared__ double scratch[32];
    __shared__ double sq[3];
    double part[3] = {0.0, 0.0, 0.0};
    for (int64_t i = threadIdx.x; i < po.total; i += blockDim.x) {
        const double v = (double)g[i];
        const int grp = i < po.wp ? 0 : (i < po.wv ? 1 : 2);
        part[grp] += v * v;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double t = block_sum(part[k], scratch);
        if (threadIdx.x == 0) sq[k] = t;
    }
    __syncthreads();
    const double total = sqrt(sq[0] + sq[1] + sq[2]);
    double coef = 1.0;      // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total + 1e-6), clamped to 1
    if (max_norm > 0.f) { coef = (double)max_norm / (total + 1e-6); coef = coef > 1.0 ? 1.0 : coef; }
    if (threadIdx.x == 0 && metrics) {
        metrics[GS_M_GRAD_NORM_ALL] = total;
        metrics[GS_M_GRAD_NORM_BACKBONE] = sqrt(sq[0]);
        metrics[GS_M_GRAD_NORM_POLICY] = sqrt(sq[1]);
        metrics[GS_M_GRAD_NORM_VALUE] = sqrt(sq[2]);
        metrics[GS_M_CLIP_COEF] = coef;
    }
    if (coef < 1.0) {
        const float c = (float)coef;
        for (int64_t i = threadIdx.x; i < po.total; i += blockDim.x) g[i] *= c;
    }
}

// ---- Adam, single-tensor math of torch.optim.Adam (no amsgrad / weight decay / maximize) ---------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, int64_t* __restrict__ step_count, float lr, float beta1, float beta2, float eps) {
    const int64_t step = *step_count + 1;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}
__global__ void adam_bump_kernel(int64_t* step_count) { *step_count += 1; }

// ---- host launchers ----------------------------------------------------------------------------------------------------
static MlpDev to_dev(const gs_mlp_t* m) {
    MlpDev d;
    d.D = m->obs_dim; d.H1 = m->hidden1; d.H2 = m->hidden2; d.A = m->n_actions; d.has_value = m->has_value; d.act = m->activation;
    d.w1 = m->w1; d.b1 = m->b1; d.w2 = m->w2; d.b2 = m->b2; d.wp = m->wp; d.bp = m->bp; d.wv = m->wv; d.bv = m->bv;
    return d;
}

int validate_mlp(const gs_mlp_t* m) {
    if (!m) GS_FAIL("mlp is NULL");
    if (m->obs_dim < 1 || m->obs_dim > kDP) GS_FAIL("obs_dim %d unsupported (1..%d)", m->obs_dim, kDP);
    if (m->n_actions < 2 || m->n_actions + (m->has_value ? 1 : 0) > kNH) GS_FAIL("n_actions %d unsupported (2..3)", m->n_actions);
    const int h1 = m->hidden1, h2 = m->hidden2;
    const bool ok = (h1 == 64 && h2 == 0) || (h1 == 64 && h2 == 64) || (h1 == 128 && h2 == 128) || (h1 == 256 && h2 == 256);
    if (!ok) GS_FAIL("hidden dims (%d,%d) unsupported by the engine: (64,), (64,64), (128,128), (256,256)", h1, h2);
    if (m->activation != GS_ACT_RELU && m->activation != GS_ACT_TANH) GS_FAIL("activation %d unsupported", m->activation);
    if (!m->w1 || !m->b1 || !m->wp || !m->bp || (h2 > 0 && (!m->w2 || !m->b2)) || (m->has_value && (!m->wv || !m->bv)))
        GS_FAIL("mlp has NULL weight pointers");
    return 0;
}

struct UpdateWs {  // workspace carve-up
    float* grad_partials;
    double* metric_partials;
    uint32_t* dead;
    double* sq;
    uint32_t* offs;   // time-major offset of every minibatch sample (gather_offsets_kernel), max_batch entries
};

static int update_grid(int device) { return 2 * sm_count(device); }
int64_t wide_scratch_bytes(int64_t max_batch);   // update_wide.cu

static int64_t ws_bytes(const gs_mlp_t* m, int device, int64_t max_batch) {
    const int64_t P = mlp_param_count(m->obs_dim, m->hidden1, m->hidden2, m->n_actions, m->has_value);
    const int grid = update_grid(device);
    int64_t b = 0;
    b += ((int64_t)grid * ((P + 3) & ~3ll) * 4 + 255) / 256 * 256;
    b += ((int64_t)grid * PM_N * 8 + 255) / 256 * 256;
    b += ((int64_t)(m->hidden1 + m->hidden2) * 4 + 255) / 256 * 256;
    b += 256;
    b += (max_batch * 4 + 255) / 256 * 256;
    if (m->hidden1 == 256 && m->hidden2 == 256) b += wide_scratch_bytes(max_batch);   // staged W2 + the h1 / dz2 tiles of update_wide.cu
    return b;
}

static UpdateWs carve(void* ws, const gs_mlp_t* m, int grid) {
    const int64_t P = mlp_param_count(m->obs_dim, m->hidden1, m->hidden2, m->n_actions, m->has_value);
    char* p = (char*)ws;
    UpdateWs w;
    w.grad_partials = (float*)p; p += ((int64_t)grid * ((P + 3) & ~3ll) * 4 + 255) / 256 * 256;
    w.metric_partials = (double*)p; p += ((int64_t)grid * PM_N * 8 + 255) / 256 * 256;
    w.dead = (uint32_t*)p; p += ((int64_t)(m->hidden1 + m->hidden2) * 4 + 255) / 256 * 256;
    w.sq = (double*)p; p += 256;
    w.offs = (uint32_t*)p;
    return w;
}

// tensor-core kernels of update_f16.cu (H x H networks, H = 64 / 128, obs_dim <= 7)
template <int ALGO>
int launch_update_f16(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                     const uint32_t* offs, const void* records, float* grad_partials, int64_t pstride, double* metric_partials, uint32_t* dead,
                     int grid, cudaStream_t st);
int f16_sets(int H);
// 256 x 256 relu networks: update_wide.cu (fused kernel + layer-2 weight-gradient kernel over tiles stored in the workspace)
template <int ALGO>
int launch_update_wide(const MlpDev& md, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom, const double* ret_mom,
                       const uint32_t* offs, const void* records, void* scratch, float* grad_partials, int64_t pstride, double* metric_partials,
                       uint32_t* dead, int grid, cudaStream_t st);
int launch_rollout_pack(const BatchDev& b, void* packed, int device, cudaStream_t st);
int launch_batch_pack(const BatchDev& b, const uint32_t* offs, void* packed, cudaStream_t st);

// does the tensor-core kernel serve this network / rollout?  (everything else runs update_kernel on the fp32 FMA pipe)
static int update_impl();
static bool tensor_path_for(int h1, int h2, int act, int obs_dim, int64_t T, int64_t N) {
    const bool width = h1 == 64 || h1 == 128 || (h1 == 256 && act == GS_ACT_RELU);
    return h1 == h2 && width && obs_dim <= 7 && update_impl() == 0 && T * N < (1ll << 32);
}

// Sample-id translation of the whole minibatch in one pass: offs[pos] = time-major offset of minibatch position pos.  The
// keyed Feistel walk + env-major -> time-major division is a ~1000-instruction dependent chain per sample (cycle walking
// diverges inside a warp); done here it is spread over the whole GPU instead of sitting on the update kernel's critical path.
// With f0 / f1 given it also accumulates {sum, sum of squares, count} of those rollout fields over the minibatch (the batch
// normalisation moments of gs_batch_moments) into mom[0..2] / mom[3..5]: one pass instead of two over the permutation.
__global__ void gather_offsets_kernel(BatchDev b, uint32_t* __restrict__ offs, const float* __restrict__ f0, const float* __restrict__ f1,
                                      double* __restrict__ mom) {
    __shared__ double scratch[32];
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t off = 0;
    if (pos < b.n) { off = sample_offset(b, pos); offs[pos] = (uint32_t)off; }
    if (f0 || f1) {
        double v0 = 0.0, v1 = 0.0;
        if (pos < b.n) { if (f0) v0 = (double)__ldg(f0 + off); if (f1) v1 = (double)__ldg(f1 + off); }
        if (f0) {
            const double s = block_sum(v0, scratch), s2 = block_sum(v0 * v0, scratch);
            if (threadIdx.x == 0) { atomicAdd(mom + 0, s); atomicAdd(mom + 1, s2); if (blockIdx.x == 0) atomicAdd(mom + 2, (double)b.n); }
        }
        if (f1) {
            const double s = block_sum(v1, scratch), s2 = block_sum(v1 * v1, scratch);
            if (threadIdx.x == 0) { atomicAdd(mom + 3, s); atomicAdd(mom + 4, s2); if (blockIdx.x == 0) atomicAdd(mom + 5, (double)b.n); }
        }
    }
}

// 0 = tensor cores where a kernel exists (64x64, 128x128, 256x256 relu), 1 = fp32 SIMT everywhere.  GS_UPDATE_IMPL=simt|tc overrides at load.
static int g_update_impl = -1;
static int update_impl() {
    if (g_update_impl < 0) {
        const char* e = getenv("GS_UPDATE_IMPL");
        g_update_impl = (e && strcmp(e, "simt") == 0) ? 1 : 0;
    }
    return g_update_impl;
}

template <class C, int ALGO>
static int launch_update_cfg(const gs_mlp_t* m, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom,
                             const double* ret_mom, float* grads_flat, double* metrics, void* ws, int device, cudaStream_t st) {
    const int grid_max = update_grid(device);
    const int64_t n_tiles = (b.n + C::S - 1) / C::S;
    const int grid = (int)(n_tiles < grid_max ? n_tiles : grid_max);
    const UpdateWs w = carve(ws, m, grid_max);
    const int64_t P = mlp_param_count(m->obs_dim, m->hidden1, m->hidden2, m->n_actions, m->has_value);
    const size_t smem = (size_t)C::kSmemFloats * sizeof(float);
    const int64_t pstride = (P + 3) & ~3ll;   // 16-byte aligned partial vectors
    const MlpDev md = to_dev(m);
    int n_partials = grid, n_metric_partials = grid;
    // batch-normalisation moments the caller did not supply are taken over this minibatch, in the gather pass where there is one
    const float* mf0 = (hp.normalize_adv && !adv_mom) ? b.adv : nullptr;
    const float* mf1 = (hp.normalize_ret && !ret_mom) ? b.ret : nullptr;
    // dead-unit counters and the moment scratch are adjacent in the workspace: one memset node covers both
    // (+ the ticket counter of gs_update_finish behind the moments)
    // defer_reduce == 2: the caller guarantees a clean workspace (zero-initialised once; every finishing kernel leaves the dead-unit
    // counters and the ticket zeroed), so no memset node sits between the previous step tail and this update kernel -- with one in between
    // the programmatic dependent launch does not pair the two kernels and the update kernel's prologue runs after the tail, not under it
    const bool clean = b.defer_reduce == 2 && !mf0 && !mf1;
    if (!clean && (track || mf0 || mf1 || b.defer_reduce)) GS_CUDA(cudaMemsetAsync(w.dead, 0, (size_t)((char*)w.sq - (char*)w.dead) + 64, st));
    if (mf0) adv_mom = w.sq;
    if (mf1) ret_mom = w.sq + 3;
    const bool tensor_path = tensor_path_for(C::H1, C::H2, md.act, md.D, b.T, b.N);
    if (!C::kPersist && !tensor_path) GS_CUDA(cudaMemsetAsync(grads_flat, 0, (size_t)P * 4, st));
    if (!tensor_path && (mf0 || mf1)) {
        int64_t blocks = (b.n + 255) / 256;
        const int cap = 4 * sm_count(device);
        if (blocks > cap) blocks = cap;
        if (mf0) { batch_moments_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, mf0, w.sq); GS_LAUNCH_CHECK(); }
        if (mf1) { batch_moments_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, mf1, w.sq + 3); GS_LAUNCH_CHECK(); }
    }
    if (tensor_path) {
        uint32_t* offs = b.offsets ? b.offsets : w.offs;
        if (!b.prepared || mf0 || mf1) {
            gather_offsets_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, st>>>(b, offs, mf0, mf1, w.sq);
            GS_LAUNCH_CHECK();
        }
        const int64_t tiles128 = (b.n + 127) / 128;
        const int sms = sm_count(device);
        const int ctas = (int)(tiles128 < sms ? tiles128 : sms);
        n_partials = ctas * f16_sets(C::H1);
        n_metric_partials = ctas;
        const void* records = b.packed;
        const uint32_t* koffs = offs;
        void* scratch = nullptr;
        if (!records) {     // no rollout records (gs_rollout_pack): build this minibatch's records in minibatch order (stream-ordered scratch)
            GS_CUDA(cudaMallocAsync(&scratch, (size_t)b.n * GS_RECORD_FLOATS * 4, st));
            if (launch_batch_pack(b, offs, scratch, st)) return -1;
            records = scratch;
            koffs = nullptr;
        }
        int rc;
        if (C::H1 == 256) {   // staged W2 + stored tiles live behind this minibatch's offsets (ws_bytes(b.n) covers them)
            void* wide = (char*)w.offs + (b.n * 4 + 255) / 256 * 256;
            rc = launch_update_wide<ALGO>(md, b, hp, track, adv_mom, ret_mom, koffs, records, wide, w.grad_partials, pstride, w.metric_partials, w.dead, ctas, st);
        } else {
            rc = launch_update_f16<ALGO>(md, b, hp, track, adv_mom, ret_mom, koffs, records, w.grad_partials, pstride, w.metric_partials, w.dead, ctas, st);
        }
        if (scratch) GS_CUDA(cudaFreeAsync(scratch, st));
        if (rc) return -1;
    } else {
        auto kern = track ? update_kernel<C, ALGO, true> : update_kernel<C, ALGO, false>;
        GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreads, smem, st>>>(md, b, hp, adv_mom, ret_mom, w.grad_partials, grads_flat, w.metric_partials, w.dead, pstride);
        GS_LAUNCH_CHECK();
    }
    if (b.defer_reduce) return 0;             // gs_update_finish folds the partials (same n_partials: partials_for())
    const bool have_partials = C::kPersist || tensor_path;
    reduce_and_finalize_kernel<<<(unsigned)((P + 255) / 256) + 1, 256, 0, st>>>(have_partials ? w.grad_partials : nullptr, n_partials, P, pstride, grads_flat,
                                                                                  w.metric_partials, n_metric_partials, ALGO, m->hidden1, m->hidden2,
                                                                                  track ? 1 : 0, hp.vf_coef, hp.ent_coef, hp.normalize_adv,
                                                                                  hp.normalize_ret, w.dead, metrics);
    GS_LAUNCH_CHECK();
    if (track) GS_CUDA(cudaMemsetAsync(w.dead, 0, (size_t)((char*)w.sq - (char*)w.dead), st));   // every path leaves the counters clean
    return 0;
}

template <int ALGO>
static int launch_update(const gs_mlp_t* m, const BatchDev& b, const HpDev& hp, bool track, const double* adv_mom,
                         const double* ret_mom, float* grads_flat, double* metrics, void* ws, int64_t ws_size, cudaStream_t st) {
    if (validate_mlp(m)) return -1;
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    if (b.n <= 0) GS_FAIL("empty minibatch");
    if (ws_size < ws_bytes(m, device, b.n))
        GS_FAIL("workspace too small for a %lld-sample minibatch: %lld < %lld (gs_update_workspace_bytes)", (long long)b.n, (long long)ws_size,
                (long long)ws_bytes(m, device, b.n));
    if (b.D != m->obs_dim) GS_FAIL("batch obs_dim %d != mlp obs_dim %d", b.D, m->obs_dim);
    const int h1 = m->hidden1, h2 = m->hidden2;
#define GS_DISPATCH(H1, H2, S) \
    return launch_update_cfg<TileCfg<H1, H2, S>, ALGO>(m, b, hp, track, adv_mom, ret_mom, grads_flat, metrics, ws, device, st)
    if (h1 == 64 && h2 == 64) GS_DISPATCH(64, 64, 128);
    if (h1 == 64 && h2 == 0) GS_DISPATCH(64, 0, 128);
    if (h1 == 128 && h2 == 128) GS_DISPATCH(128, 128, 64);
    if (h1 == 256 && h2 == 256) GS_DISPATCH(256, 256, 64);
#undef GS_DISPATCH
    GS_FAIL("no kernel for hidden dims (%d,%d)", h1, h2);
}

static BatchDev to_dev(const gs_batch_t* b) {
    BatchDev d;
    d.n = b->n; d.idx = b->idx; d.perm_key = b->perm_key; d.perm_offset = b->perm_offset; d.perm_len = b->perm_len;
    d.idx_map = b->idx_map; d.T = b->T; d.D = b->obs_dim; d.N = b->N; d.obs = b->obs; d.actions = b->actions;
    d.logp_old = b->logp_old; d.values_old = b->values_old; d.adv = b->adv; d.ret = b->ret;
    d.packed = b->packed;
    d.prepared = b->prepared;
    d.defer_reduce = b->defer_reduce;
    d.offsets = b->offsets;
    d.perm_bits = b->perm_len > 0 ? feistel_bits((uint64_t)b->perm_len) : 0;
    return d;
}

}  // namespace gs

using namespace gs;

extern "C" {

int gs_set_update_impl(int impl) {
    if (impl != 0 && impl != 1) GS_FAIL("gs_set_update_impl: 0 = tensor cores where available, 1 = fp32 SIMT");
    g_update_impl = impl;
    return 0;
}

int64_t gs_mlp_param_count(const gs_mlp_t* m) {
    if (validate_mlp(m)) return -1;
    return mlp_param_count(m->obs_dim, m->hidden1, m->hidden2, m->n_actions, m->has_value);
}

int64_t gs_update_workspace_bytes(const gs_mlp_t* m, int device, int64_t max_batch) {
    if (validate_mlp(m)) return -1;
    if (max_batch < 0) GS_FAIL("gs_update_workspace_bytes: negative max_batch");
    return ws_bytes(m, device, max_batch);
}

int gs_batch_moments(const gs_batch_t* batch, const float* field, double* out, void* stream) {
    if (!batch || !field || !out) GS_FAIL("gs_batch_moments: NULL argument");
    if (batch->n <= 0) GS_FAIL("empty minibatch");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    const BatchDev b = to_dev(batch);
    int64_t blocks = (b.n + 255) / 256;
    const int cap = 4 * sm_count(device);
    if (blocks > cap) blocks = cap;
    batch_moments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, field, out);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_rollout_pack(const gs_batch_t* batch, float* packed, void* stream) {
    if (!batch || !packed) GS_FAIL("gs_rollout_pack: NULL argument");
    if (!batch->obs || !batch->actions || !batch->logp_old || !batch->adv || !batch->ret) GS_FAIL("gs_rollout_pack: batch has NULL arrays");
    if (batch->T <= 0 || batch->N <= 0 || batch->obs_dim <= 0 || batch->obs_dim > 7) GS_FAIL("gs_rollout_pack: bad shape (obs_dim 1..7)");
    if (((uintptr_t)packed & 15) != 0) GS_FAIL("gs_rollout_pack: packed must be 16-byte aligned");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    return launch_rollout_pack(to_dev(batch), packed, device, (cudaStream_t)stream);
}

int gs_batch_prepare(const gs_mlp_t* mlp, const gs_batch_t* batch, int want_adv, int want_ret, double* moments, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    if (!mlp || !batch || !moments || !workspace) GS_FAIL("gs_batch_prepare: NULL argument");
    if (validate_mlp(mlp)) return -1;
    if (batch->n <= 0) GS_FAIL("empty minibatch");
    if ((want_adv && !batch->adv) || (want_ret && !batch->ret)) GS_FAIL("gs_batch_prepare: batch has NULL arrays");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    if (workspace_bytes < ws_bytes(mlp, device, batch->n)) GS_FAIL("gs_batch_prepare: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const BatchDev b = to_dev(batch);
    GS_CUDA(cudaMemsetAsync(moments, 0, 6 * sizeof(double), st));
    const float* f0 = want_adv ? b.adv : nullptr;
    const float* f1 = want_ret ? b.ret : nullptr;
    const bool tensor_path = tensor_path_for(mlp->hidden1, mlp->hidden2, mlp->activation, mlp->obs_dim, b.T, b.N);
    if (tensor_path) {
        const UpdateWs w = carve(workspace, mlp, update_grid(device));
        gather_offsets_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, st>>>(b, b.offsets ? b.offsets : w.offs, f0, f1, moments);
        GS_LAUNCH_CHECK();
    } else {
        int64_t blocks = (b.n + 255) / 256;
        const int cap = 4 * sm_count(device);
        if (blocks > cap) blocks = cap;
        if (f0) { batch_moments_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, f0, moments); GS_LAUNCH_CHECK(); }
        if (f1) { batch_moments_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, f1, moments + 3); GS_LAUNCH_CHECK(); }
    }
    return 0;
}

int gs_ppo_step(const gs_mlp_t* mlp, const gs_batch_t* batch, const gs_ppo_hparams_t* hp, const double* adv_moments,
                float* grads_flat, double* metrics, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!mlp || !batch || !hp || !grads_flat || !metrics || !workspace) GS_FAIL("gs_ppo_step: NULL argument");
    if (!mlp->has_value) GS_FAIL("PPO requires a policy with a value head");  // agents/ppo/ppo_agent.py:41-44
    if (!batch->values_old || !batch->adv || !batch->ret || !batch->logp_old || !batch->actions || !batch->obs)
        GS_FAIL("gs_ppo_step: batch has NULL arrays");
    HpDev h;
    h.clip_lo = (float)(1.0 - (double)hp->clip_range);
    h.clip_hi = (float)(1.0 + (double)hp->clip_range);
    h.clip_vf = hp->clip_range_vf; h.vf_coef = hp->vf_coef; h.ent_coef = hp->ent_coef;
    h.normalize_adv = hp->normalize_adv; h.normalize_ret = 0; h.policy_targets = 1;
    return launch_update<ALGO_PPO>(mlp, to_dev(batch), h, hp->track_activations != 0, adv_moments, nullptr, grads_flat, metrics,
                                   workspace, workspace_bytes, (cudaStream_t)stream);
}

int gs_reinforce_step(const gs_mlp_t* mlp, const gs_batch_t* batch, const gs_reinforce_hparams_t* hp, const double* ret_moments,
                      const double* adv_moments, float* grads_flat, double* metrics, void* workspace, int64_t workspace_bytes,
                      void* stream) {
    if (!mlp || !batch || !hp || !grads_flat || !metrics || !workspace) GS_FAIL("gs_reinforce_step: NULL argument");
    if (hp->policy_targets != 0 && hp->policy_targets != 1) GS_FAIL("Invalid policy targets: %d", hp->policy_targets);
    if (!batch->adv || !batch->ret || !batch->logp_old || !batch->actions || !batch->obs) GS_FAIL("gs_reinforce_step: batch has NULL arrays");
    HpDev h;
    h.clip_lo = h.clip_hi = h.clip_vf = h.vf_coef = 0.f;
    h.ent_coef = hp->ent_coef;
    h.normalize_adv = hp->normalize_adv; h.normalize_ret = hp->normalize_returns; h.policy_targets = hp->policy_targets;
    return launch_update<ALGO_REINFORCE>(mlp, to_dev(batch), h, hp->track_activations != 0, adv_moments, ret_moments, grads_flat,
                                         metrics, workspace, workspace_bytes, (cudaStream_t)stream);
}

int gs_clip_grad_norm(const gs_mlp_t* mlp, float* grads_flat, float max_norm, double* metrics, void* stream) {
    if (validate_mlp(mlp)) return -1;
    if (!grads_flat || !metrics) GS_FAIL("gs_clip_grad_norm: NULL argument");
    const ParamOffsets po = param_offsets(mlp->obs_dim, mlp->hidden1, mlp->hidden2, mlp->n_actions, mlp->has_value);
    cudaStream_t st = (cudaStream_t)stream;
    if (po.total <= (1 << 17)) {          // every network of the registry: one block, one launch, no atomics
        clip_grad_norm_kernel<<<1, 1024, 0, st>>>(grads_flat, po, max_norm, metrics);
        GS_LAUNCH_CHECK();
        return 0;
    }
    double* sq = metrics + GS_M_SCRATCH;  // squared norms of [backbone, policy_head, value_head]
    GS_CUDA(cudaMemsetAsync(sq, 0, 3 * sizeof(double), st));
    const int blocks = (int)((po.total + 1023) / 1024 < 64 ? (po.total + 1023) / 1024 : 64);
    grad_norm_kernel<<<blocks, 256, 0, st>>>(grads_flat, po, sq);
    GS_LAUNCH_CHECK();
    clip_scale_kernel<<<blocks, 256, 0, st>>>(grads_flat, po.total, sq, max_norm, metrics);
    GS_LAUNCH_CHECK();
    return 0;
}

// ---- NVLink peer group (CUDA IPC) --------------------------------------------------------------------------------------
struct gs_peer {
    int rank = 0, world = 1, device = 0;
    int64_t stride = 0;                 // floats per slot (16-byte multiple)
    void* base = nullptr;               // own allocation: [2 phases][world][stride] 8-byte words {value bits, epoch}
    void* mapped[GS_PEER_MAX_WORLD] = {};   // peers' allocations (own rank: base)
    bool connected = false;
    uint32_t epoch = 0, mom_epoch = 0;  // call counters of gs_update_finish / gs_peer_allreduce_f64 (identical on every rank)
};
static size_t peer_slot_bytes(const gs_peer* p) { return ((size_t)2 * p->world * p->stride * 8 + 255) / 256 * 256; }
static size_t peer_mom_bytes(const gs_peer* p) { return (size_t)2 * p->world * (2 * kPeerMomDoubles) * 8; }   // behind the gradient slots (+ 256)

int gs_peer_create(int rank, int world_size, int64_t max_floats, int device, gs_peer_t** out, void* handle_out_host) {
    if (!out || !handle_out_host) GS_FAIL("gs_peer_create: NULL argument");
    if (world_size < 2 || world_size > GS_PEER_MAX_WORLD) GS_FAIL("gs_peer_create: world_size %d outside 2..%d", world_size, GS_PEER_MAX_WORLD);
    if (rank < 0 || rank >= world_size) GS_FAIL("gs_peer_create: bad rank %d/%d", rank, world_size);
    if (max_floats <= 0) GS_FAIL("gs_peer_create: max_floats must be positive");
    static_assert(sizeof(cudaIpcMemHandle_t) == GS_PEER_HANDLE_BYTES, "IPC handle size");
    GS_CUDA(cudaSetDevice(device));
    gs_peer* p = new gs_peer();
    p->rank = rank; p->world = world_size; p->device = device;
    p->stride = (max_floats + 3) & ~3ll;
    const size_t bytes = peer_slot_bytes(p) + 256 + peer_mom_bytes(p);
    if (cudaMalloc(&p->base, bytes) != cudaSuccess) { delete p; GS_FAIL("gs_peer_create: cudaMalloc of %zu bytes failed", bytes); }
    cudaMemset(p->base, 0, bytes);
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p->base);
    if (e != cudaSuccess) { cudaFree(p->base); delete p; GS_FAIL("gs_peer_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle_out_host, &h, sizeof(h));
    *out = p;
    return 0;
}

int gs_peer_connect(gs_peer_t* p, const void* all_handles_host) {
    if (!p || !all_handles_host) GS_FAIL("gs_peer_connect: NULL argument");
    if (p->connected) GS_FAIL("gs_peer_connect: already connected");
    GS_CUDA(cudaSetDevice(p->device));
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) { p->mapped[r] = p->base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles_host + (size_t)r * GS_PEER_HANDLE_BYTES, sizeof(h));
        const cudaError_t e = cudaIpcOpenMemHandle(&p->mapped[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (q != p->rank && p->mapped[q]) { cudaIpcCloseMemHandle(p->mapped[q]); p->mapped[q] = nullptr; }
            GS_FAIL("gs_peer_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
    }
    p->connected = true;
    return 0;
}

int gs_peer_destroy(gs_peer_t* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; ++r) if (r != p->rank && p->mapped[r]) cudaIpcCloseMemHandle(p->mapped[r]);
    if (p->base) cudaFree(p->base);
    delete p;
    return 0;
}

int gs_peer_allreduce_f64(gs_peer_t* peer, double* data, int64_t n, void* stream) {
    if (!peer || !data) GS_FAIL("gs_peer_allreduce_f64: NULL argument");
    if (!peer->connected) GS_FAIL("gs_peer_allreduce_f64: peer group is not connected (gs_peer_connect)");
    if (n <= 0 || n > kPeerMomDoubles) GS_FAIL("gs_peer_allreduce_f64: n = %lld outside 1..%lld", (long long)n, (long long)kPeerMomDoubles);
    PeerMomDev pm = {};
    pm.rank = peer->rank; pm.world = peer->world; pm.epoch = ++peer->mom_epoch;
    for (int r = 0; r < peer->world; ++r) pm.slots[r] = (unsigned long long*)((char*)peer->mapped[r] + peer_slot_bytes(peer) + 256);
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pm, data, (int)n);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_update_finish(const gs_mlp_t* mlp, const gs_batch_t* batch, const gs_finish_t* fin, float* grads_flat, const gs_adam_t* adam, gs_peer_t* peer,
                     double* metrics, double* metrics_sum, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!mlp || !batch || !fin || !grads_flat || !metrics || !workspace) GS_FAIL("gs_update_finish: NULL argument");
    if (validate_mlp(mlp)) return -1;
    if (batch->n <= 0) GS_FAIL("gs_update_finish: batch must be the deferred step's minibatch");
    if (fin->algo != ALGO_PPO && fin->algo != ALGO_REINFORCE) GS_FAIL("gs_update_finish: bad algo %d", fin->algo);
    if (adam && (!adam->params_flat || !adam->exp_avg || !adam->exp_avg_sq || !adam->step_count)) GS_FAIL("gs_update_finish: adam has NULL pointers");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    if (workspace_bytes < ws_bytes(mlp, device, 0)) GS_FAIL("gs_update_finish: workspace too small");
    const int64_t P = mlp_param_count(mlp->obs_dim, mlp->hidden1, mlp->hidden2, mlp->n_actions, mlp->has_value);
    const UpdateWs w = carve(workspace, mlp, update_grid(device));
    FinishDev f;
    // the partial vectors the deferred step left behind (mirrors launch_update_cfg)
    const bool persist = mlp->hidden1 <= 64 && mlp->hidden2 <= 64;
    const int S = (mlp->hidden1 <= 64) ? 128 : 64;
    int64_t n_tiles = (batch->n + S - 1) / S;
    const int grid_max = update_grid(device);
    int n_cta = (int)(n_tiles < grid_max ? n_tiles : grid_max);
    int n_metric_cta = n_cta;
    const bool tensor_path = tensor_path_for(mlp->hidden1, mlp->hidden2, mlp->activation, mlp->obs_dim, batch->T, batch->N);
    if (tensor_path) {
        const int sms = sm_count(device);
        n_tiles = (batch->n + 127) / 128;
        n_metric_cta = (int)(n_tiles < sms ? n_tiles : sms);
        n_cta = n_metric_cta * f16_sets(mlp->hidden1);
    }
    f.partials = (persist || tensor_path) ? w.grad_partials : nullptr;
    f.n_cta = n_cta; f.n_metric_cta = n_metric_cta;
    f.P = P; f.pstride = (P + 3) & ~3ll;
    f.metric_partials = w.metric_partials;
    f.algo = fin->algo; f.H1 = mlp->hidden1; f.H2 = mlp->hidden2; f.track = fin->track_activations ? 1 : 0;
    f.normalize_adv = fin->normalize_adv; f.normalize_ret = fin->normalize_ret;
    f.vf_coef = fin->vf_coef; f.ent_coef = fin->ent_coef; f.max_norm = fin->max_grad_norm;
    f.dead = w.dead; f.ticket = (uint32_t*)((char*)w.sq + 56);
    AdamDev ad = {};
    if (adam) { ad.p = adam->params_flat; ad.m = adam->exp_avg; ad.v = adam->exp_avg_sq; ad.step = adam->step_count;
                ad.lr = adam->lr; ad.beta1 = adam->beta1; ad.beta2 = adam->beta2; ad.eps = adam->eps; }
    PeerDev pd = {};
    pd.world = 1;
    if (peer) {
        if (!peer->connected) GS_FAIL("gs_update_finish: peer group is not connected (gs_peer_connect)");
        if (peer->device != device) GS_FAIL("gs_update_finish: peer group belongs to device %d, current device is %d", peer->device, device);
        if (f.pstride > peer->stride) GS_FAIL("gs_update_finish: peer slots hold %lld floats, the model has %lld", (long long)peer->stride, (long long)P);
        pd.rank = peer->rank; pd.world = peer->world; pd.stride = peer->stride;
        pd.epoch = ++peer->epoch;
        for (int r = 0; r < peer->world; ++r) pd.slots[r] = (unsigned long long*)peer->mapped[r];
    }
    const ParamOffsets po = param_offsets(mlp->obs_dim, mlp->hidden1, mlp->hidden2, mlp->n_actions, mlp->has_value);
    f.split = P > kFinishSplitP ? 1 : 0;
    const unsigned blocks = (unsigned)((P + 63) / 64) + (f.split ? 0u : 1u);     // + the metrics block
    update_finish_kernel<<<blocks, kFinishThreads, 0, (cudaStream_t)stream>>>(f, po, grads_flat, ad, pd, metrics, metrics_sum);
    GS_LAUNCH_CHECK();
    if (f.split) {
        // the partial vectors have been consumed: their memory holds the squared-norm partials (the next update kernel writes partials only
        // after its griddepcontrol.wait, i.e. after finish_apply_kernel has completed)
        double* sq_part = reinterpret_cast<double*>(w.grad_partials);
        const unsigned nb = (unsigned)((P + kFinishThreads - 1) / kFinishThreads);
        finish_receive_kernel<<<nb, kFinishThreads, 0, (cudaStream_t)stream>>>(f, po, grads_flat, ad, pd, sq_part);
        GS_LAUNCH_CHECK();
        finish_apply_kernel<<<nb, kFinishThreads, 0, (cudaStream_t)stream>>>(f, po, grads_flat, ad, sq_part, metrics, metrics_sum);
        GS_LAUNCH_CHECK();
    }
    return 0;
}

int gs_adam_step(float* params_flat, const float* grads_flat, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step_count,
                 float lr, float beta1, float beta2, float eps, void* stream) {
    if (!params_flat || !grads_flat || !exp_avg || !exp_avg_sq || !step_count || n <= 0) GS_FAIL("gs_adam_step: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)((n + 255) / 256 < 296 ? (n + 255) / 256 : 296);
    adam_kernel<<<blocks, 256, 0, st>>>(params_flat, grads_flat, exp_avg, exp_avg_sq, n, step_count, lr, beta1, beta2, eps);
    GS_LAUNCH_CHECK();
    adam_bump_kernel<<<1, 1, 0, st>>>(step_count);
    GS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

#ifdef GS_FINISH_TRACE
extern "C" int gs_debug_finish_trace(unsigned long long* host_out /* [8] */) {
    return (int)cudaMemcpyFromSymbol(host_out, gs::g_fin_trace, sizeof(unsigned long long) * 8);
}
#endif
