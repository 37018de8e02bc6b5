// mlp_tile.cuh — CTA-tile MLP forward / backward on fp32 FMA pipes, shared by gs_policy_act, the fused collect
// kernel and the PPO / REINFORCE update kernels.
//
// Replaces MLPActorCritic.forward / MLPPolicy.forward (utils/models.py:233-346) and the autograd graph behind
// PPOAgent.losses_for_batch (agents/ppo/ppo_agent.py:21-152).
//
// Layout: a CTA of 256 threads owns a tile of S samples.  Activations live in shared memory, sample-major with a
// 4-float pad per row ([S][H+4]: the pad makes the 16 B row loads of 8 consecutive rows hit 8 distinct bank groups).
// Weights are torch nn.Linear layout [out][in]; a 64x64 block is staged into wbuf[64][68] (resident for the whole
// kernel when the layer IS 64x64).  Register tiles: forward / dgrad TS x 4 outputs per thread fed by 16 B shared loads
// (10.7 FMA per LDS.128), wgrad 8 x 4 per thread kept in registers across every tile of the CTA.
#pragma once

#include "common.cuh"

namespace gs {

constexpr int kThreads = 256;
constexpr int kDP = 8;        // observation dim padded to 8 (D <= 8)
constexpr int kLDX = 12;      // row stride of the padded observation tile
constexpr int kNH = 4;        // head rows: A policy logits (A <= 3) + 1 value row
constexpr int kWLD = 68;      // row stride of a staged 64x64 weight block

struct MlpDev {  // device copy of gs_mlp_t (passed by value to kernels)
    int D, H1, H2, A, has_value, act;
    const float *w1, *b1, *w2, *b2, *wp, *bp, *wv, *bv;
};

__host__ __device__ inline int64_t mlp_param_count(int D, int H1, int H2, int A, int has_value) {
    const int HL = H2 > 0 ? H2 : H1;
    int64_t p = (int64_t)H1 * D + H1;
    if (H2 > 0) p += (int64_t)H2 * H1 + H2;
    p += (int64_t)A * HL + A;
    if (has_value) p += HL + 1;
    return p;
}

// offsets of each parameter tensor inside the flat vector (parameters() order)
struct ParamOffsets {
    int64_t w1, b1, w2, b2, wp, bp, wv, bv, total;
};
__host__ __device__ inline ParamOffsets param_offsets(int D, int H1, int H2, int A, int has_value) {
    ParamOffsets o;
    const int HL = H2 > 0 ? H2 : H1;
    int64_t p = 0;
    o.w1 = p; p += (int64_t)H1 * D;
    o.b1 = p; p += H1;
    o.w2 = p; if (H2 > 0) p += (int64_t)H2 * H1;
    o.b2 = p; if (H2 > 0) p += H2;
    o.wp = p; p += (int64_t)A * HL;
    o.bp = p; p += A;
    o.wv = p; if (has_value) p += HL;
    o.bv = p; if (has_value) p += 1;
    o.total = p;
    return o;
}

template <int H1_, int H2_, int S_>
struct TileCfg {
    static constexpr int H1 = H1_, H2 = H2_, S = S_;
    static constexpr int HL = H2 > 0 ? H2 : H1;   // width feeding the heads
    static constexpr int TS = S / 16;             // samples per thread in a forward / dgrad register tile
    static constexpr int LD1 = H1 + 4, LD2 = (H2 > 0 ? H2 : 0) + 4, LDL = HL + 4;
    static constexpr bool kResidentW2 = (H1 == 64 && H2 == 64);  // the 64x64 layer never leaves shared memory
    static constexpr bool kPersist = (H1 <= 64 && H2 <= 64);     // wgrad accumulators stay in registers all kernel long
    static_assert(S % 16 == 0 && S <= kThreads, "tile size");
    static_assert(H1 % 64 == 0 && (H2 == 0 || H2 % 64 == 0), "hidden widths are multiples of 64");
    // shared memory carve-up (floats)
    static constexpr int oXS = 0;                          // [S][12]   observations
    static constexpr int oA1 = oXS + S * kLDX;             // [S][LD1]  h1, later dZ1
    static constexpr int oA2 = oA1 + S * LD1;              // [S][LD2]  h2, later dZ2
    static constexpr int oWB = oA2 + (H2 > 0 ? S * LD2 : 0);  // [64][68] staged / resident 64x64 weight block
    static constexpr int oW1 = oWB + 64 * kWLD;            // [H1][12]  layer-1 weights (resident)
    static constexpr int oWH = oW1 + H1 * kLDX;            // [4][LDL]  head rows (resident)
    static constexpr int oB1 = oWH + kNH * LDL;            // [H1]
    static constexpr int oB2 = oB1 + H1;                   // [H2]
    static constexpr int oBH = oB2 + (H2 > 0 ? H2 : 0);    // [4]
    static constexpr int oG = oBH + kNH;                   // [S][4]    dLoss/d(head outputs)
    static constexpr int oEnd = oG + S * kNH;
    static constexpr int kSmemFloats = oEnd;
};

__device__ __forceinline__ float act_fwd(float z, int act) { return act == GS_ACT_RELU ? fmaxf(z, 0.0f) : tanhf(z); }
// derivative from the activation OUTPUT h (relu: h > 0, tanh: 1 - h^2)
__device__ __forceinline__ float act_bwd(float h, int act) { return act == GS_ACT_RELU ? (h > 0.0f ? 1.0f : 0.0f) : 1.0f - h * h; }

// ---- weight staging ---------------------------------------------------------------------------------------
// 64x64 block (rows n0.., cols k0..) of a row-major [N][K] global matrix -> wbuf[64][68]
__device__ __forceinline__ void stage_w64(float* wbuf, const float* __restrict__ Wg, int K, int n0, int k0) {
    for (int i = threadIdx.x; i < 64 * 16; i += kThreads) {
        const int r = i >> 4, c4 = i & 15;
        const float4 v = __ldg(reinterpret_cast<const float4*>(Wg + (size_t)(n0 + r) * K + k0) + c4);
        *reinterpret_cast<float4*>(wbuf + r * kWLD + 4 * c4) = v;
    }
}

// resident operands: layer-1 weights (zero padded to 8 inputs), head rows, biases
template <class C>
__device__ __forceinline__ void load_resident(float* sm, const MlpDev& m) {
    for (int i = threadIdx.x; i < C::H1 * kLDX; i += kThreads) {
        const int n = i / kLDX, d = i % kLDX;
        sm[C::oW1 + i] = d < m.D ? __ldg(m.w1 + n * m.D + d) : 0.0f;
    }
    for (int i = threadIdx.x; i < kNH * C::LDL; i += kThreads) {
        const int r = i / C::LDL, k = i % C::LDL;
        float v = 0.0f;
        if (k < C::HL) {
            if (r < m.A) v = __ldg(m.wp + r * C::HL + k);
            else if (r == m.A && m.has_value) v = __ldg(m.wv + k);
        }
        sm[C::oWH + i] = v;
    }
    for (int i = threadIdx.x; i < C::H1; i += kThreads) sm[C::oB1 + i] = __ldg(m.b1 + i);
    if (C::H2 > 0)
        for (int i = threadIdx.x; i < C::H2; i += kThreads) sm[C::oB2 + i] = __ldg(m.b2 + i);
    if (threadIdx.x < kNH) {
        const int r = threadIdx.x;
        sm[C::oBH + r] = r < m.A ? __ldg(m.bp + r) : ((r == m.A && m.has_value) ? __ldg(m.bv) : 0.0f);
    }
    if (C::kResidentW2) stage_w64(sm + C::oWB, m.w2, C::H1, 0, 0);
}

// ---- register-tile GEMM pieces -------------------------------------------------------------------------------
// acc[j][i] += sum_{k<KC} A[(ty*TS+j)][k] * W[(tx+16i)][k]      (A: smem rows stride lda, W: smem rows stride ldw)
template <int TS, int KC>
__device__ __forceinline__ void tile_nt(float (&acc)[TS][4], const float* A, int lda, const float* W, int ldw, int tx, int ty) {
#pragma unroll 4
    for (int k = 0; k < KC; k += 4) {
        float4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(W + (tx + 16 * i) * ldw + k);
#pragma unroll
        for (int j = 0; j < TS; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(A + (ty * TS + j) * lda + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[j][i] = fmaf(a.x, w[i].x, acc[j][i]);
                acc[j][i] = fmaf(a.y, w[i].y, acc[j][i]);
                acc[j][i] = fmaf(a.z, w[i].z, acc[j][i]);
                acc[j][i] = fmaf(a.w, w[i].w, acc[j][i]);
            }
        }
    }
}

// acc[j][c] += sum_{n<64} G[(ty*TS+j)][n] * W[n][4tx+c]         (dgrad: reduce over the layer's outputs)
template <int TS>
__device__ __forceinline__ void tile_nn(float (&acc)[TS][4], const float* G, int ldg, const float* W, int ldw, int tx, int ty) {
#pragma unroll 4
    for (int n = 0; n < 64; n += 4) {
        float4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(W + (n + i) * ldw + 4 * tx);
#pragma unroll
        for (int j = 0; j < TS; ++j) {
            const float4 g = *reinterpret_cast<const float4*>(G + (ty * TS + j) * ldg + n);
            acc[j][0] = fmaf(g.x, w[0].x, acc[j][0]); acc[j][1] = fmaf(g.x, w[0].y, acc[j][1]);
            acc[j][2] = fmaf(g.x, w[0].z, acc[j][2]); acc[j][3] = fmaf(g.x, w[0].w, acc[j][3]);
            acc[j][0] = fmaf(g.y, w[1].x, acc[j][0]); acc[j][1] = fmaf(g.y, w[1].y, acc[j][1]);
            acc[j][2] = fmaf(g.y, w[1].z, acc[j][2]); acc[j][3] = fmaf(g.y, w[1].w, acc[j][3]);
            acc[j][0] = fmaf(g.z, w[2].x, acc[j][0]); acc[j][1] = fmaf(g.z, w[2].y, acc[j][1]);
            acc[j][2] = fmaf(g.z, w[2].z, acc[j][2]); acc[j][3] = fmaf(g.z, w[2].w, acc[j][3]);
            acc[j][0] = fmaf(g.w, w[3].x, acc[j][0]); acc[j][1] = fmaf(g.w, w[3].y, acc[j][1]);
            acc[j][2] = fmaf(g.w, w[3].z, acc[j][2]); acc[j][3] = fmaf(g.w, w[3].w, acc[j][3]);
        }
    }
}

// acc[i][c] += sum_{s = half, half+2, ... < S} G[s][n0 + 8tn + i] * X[s][k0 + 4tk + c]      (wgrad 64x64 block)
template <int S>
__device__ __forceinline__ void tile_tn(float (&acc)[8][4], const float* G, int ldg, const float* X, int ldx, int tn, int tk, int half) {
#pragma unroll 2
    for (int s = half; s < S; s += 2) {
        const float4 g0 = *reinterpret_cast<const float4*>(G + s * ldg + 8 * tn);
        const float4 g1 = *reinterpret_cast<const float4*>(G + s * ldg + 8 * tn + 4);
        const float4 x = *reinterpret_cast<const float4*>(X + s * ldx + 4 * tk);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(g[i], x.x, acc[i][0]);
            acc[i][1] = fmaf(g[i], x.y, acc[i][1]);
            acc[i][2] = fmaf(g[i], x.z, acc[i][2]);
            acc[i][3] = fmaf(g[i], x.w, acc[i][3]);
        }
    }
}

// ---- activation statistics of the hooked Linear outputs (utils/models.py:121-146) --------------------------------
struct ActStats {
    float sum[2], sumsq[2];   // per-thread partials of z and z^2 for backbone.0 / backbone.2
    uint32_t* dead;           // per-CTA shared-memory [H1 + H2] counters of |z| < 1e-6, flushed once to the global counters
                              // (integer atomics: deterministic; never one global atomic per sample: a unit that is dead for
                              // the whole batch would serialise them on a single address)
};

// ---- forward -------------------------------------------------------------------------------------------------
// Layer 1: a1 = act(xs @ w1^T + b1), K = 8 (zero padded), weights resident in shared memory.
template <class C, bool TRACK>
__device__ __forceinline__ void forward_layer1(float* sm, int act, int valid_rows, ActStats* st) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 1
    for (int nc = 0; nc < C::H1 / 64; ++nc) {
        float acc[C::TS][4];
#pragma unroll
        for (int j = 0; j < C::TS; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
        tile_nt<C::TS, kDP>(acc, sm + C::oXS, kLDX, sm + C::oW1 + nc * 64 * kLDX, kLDX, tx, ty);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = nc * 64 + tx + 16 * i;
            const float b = sm[C::oB1 + n];
            uint32_t n_dead = 0;
#pragma unroll
            for (int j = 0; j < C::TS; ++j) {
                const int s = ty * C::TS + j;
                const float z = acc[j][i] + b;
                if (TRACK && s < valid_rows) {
                    st->sum[0] += z; st->sumsq[0] = fmaf(z, z, st->sumsq[0]);
                    n_dead += fabsf(z) < 1e-6f ? 1u : 0u;
                }
                sm[C::oA1 + s * C::LD1 + n] = act_fwd(z, act);
            }
            if (TRACK && n_dead) atomicAdd(st->dead + n, n_dead);
        }
    }
}

// Layer 2: a2 = act(a1 @ w2^T + b2); 64x64 blocks of w2 staged through wbuf unless resident.
template <class C, bool TRACK>
__device__ __forceinline__ void forward_layer2(float* sm, const MlpDev& m, int valid_rows, ActStats* st) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 1
    for (int nc = 0; nc < C::H2 / 64; ++nc) {
        float acc[C::TS][4];
#pragma unroll
        for (int j = 0; j < C::TS; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
#pragma unroll 1
        for (int kc = 0; kc < C::H1 / 64; ++kc) {
            if (!C::kResidentW2) {
                __syncthreads();
                stage_w64(sm + C::oWB, m.w2, C::H1, nc * 64, kc * 64);
                __syncthreads();
            }
            tile_nt<C::TS, 64>(acc, sm + C::oA1 + kc * 64, C::LD1, sm + C::oWB, kWLD, tx, ty);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = nc * 64 + tx + 16 * i;
            const float b = sm[C::oB2 + n];
            uint32_t n_dead = 0;
#pragma unroll
            for (int j = 0; j < C::TS; ++j) {
                const int s = ty * C::TS + j;
                const float z = acc[j][i] + b;
                if (TRACK && s < valid_rows) {
                    st->sum[1] += z; st->sumsq[1] = fmaf(z, z, st->sumsq[1]);
                    n_dead += fabsf(z) < 1e-6f ? 1u : 0u;
                }
                sm[C::oA2 + s * C::LD2 + n] = act_fwd(z, m.act);
            }
            if (TRACK && n_dead) atomicAdd(st->dead + C::H1 + n, n_dead);
        }
    }
}

// backbone forward for the tile whose observations are in xs; ends with a barrier so every row of the last hidden
// layer is visible.  Caller must have synchronised after writing xs.
template <class C, bool TRACK>
__device__ __forceinline__ void forward_backbone(float* sm, const MlpDev& m, int valid_rows, ActStats* st) {
    forward_layer1<C, TRACK>(sm, m.act, valid_rows, st);
    __syncthreads();
    if (C::H2 > 0) {
        forward_layer2<C, TRACK>(sm, m, valid_rows, st);
        __syncthreads();
    }
}

// heads for sample row s (one thread per sample): out[r] = bh[r] + <hL[s], wh[r]>, r < 4
template <class C>
__device__ __forceinline__ void forward_heads(const float* sm, int s, float (&out)[kNH]) {
    const float* h = sm + (C::H2 > 0 ? C::oA2 : C::oA1) + s * C::LDL;
    const float* wh = sm + C::oWH;
#pragma unroll
    for (int r = 0; r < kNH; ++r) out[r] = sm[C::oBH + r];
#pragma unroll 4
    for (int k = 0; k < C::HL; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(h + k);
#pragma unroll
        for (int r = 0; r < kNH; ++r) {
            const float4 w = *reinterpret_cast<const float4*>(wh + r * C::LDL + k);
            out[r] = fmaf(a.x, w.x, out[r]); out[r] = fmaf(a.y, w.y, out[r]);
            out[r] = fmaf(a.z, w.z, out[r]); out[r] = fmaf(a.w, w.w, out[r]);
        }
    }
}

// Transcendentals of the per-sample loss: full-precision libm-style by default; a translation unit may opt into the SFU
// approximations (ex2.approx / lg2.approx based, ~2^-21 relative) by defining GS_FAST_TRANSCENDENTALS before this header.
#ifdef GS_FAST_TRANSCENDENTALS
#define GS_EXPF(x) __expf(x)
#define GS_LOGF(x) __logf(x)
#else
#define GS_EXPF(x) expf(x)
#define GS_LOGF(x) logf(x)
#endif

// log-softmax over A <= 3 logits exactly as torch: x - (max + log(sum(exp(x - max))))
__device__ __forceinline__ void log_softmax(const float* logits, int A, float* logp) {
    float mx = logits[0];
#pragma unroll
    for (int k = 1; k < 3; ++k)
        if (k < A) mx = fmaxf(mx, logits[k]);
    float se = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (k < A) se += GS_EXPF(logits[k] - mx);
    const float lse = mx + GS_LOGF(se);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (k < A) logp[k] = logits[k] - lse;
}

}  // namespace gs
