// returns_kernels.cu — GAE(lambda) and Monte-Carlo return reverse scans, episode conversion, valid-mask / index map,
// masked moments and normalisation over the time-major (T, N) rollout buffer.
//
// Replaces utils/returns_advantages.py (compute_batched_gae_advantages_and_returns :115-155, compute_batched_mc_returns
// :67-91, convert_returns_to_full_episode :93-113, _build_valid_mask_and_index_map :33-52,
// _build_idx_map_from_valid_mask :19-30, _normalize_* :55-64) and the RunningStats feeds of
// utils/rollout_collector.py:415-455.
//
// HBM-bound: one thread per env walks t = T-1..0; element (t, n) sits at t*N + n so every warp load is one fully
// coalesced line.  The recurrence is sequential in t by construction (and evaluated with the reference's exact fp32
// rounding sequence: results are bit-identical to numpy), so bandwidth comes from keeping UNROLL timesteps of loads in
// flight per thread; 64-thread CTAs keep the per-SM CTA count balanced (N/64 CTAs over 148 SMs).
#include "common.cuh"

namespace gs {

constexpr int kScanThreads = 64;
constexpr int kUnroll = 8;

// BOOT: 0 = no bootstrap values (next value of a truncated step = the following step's value, as when the reference is given None),
//       1 = dense (T, N) array, 2 = the array is identically ZERO and is not read (4 of the 22 B per element): what the collector
//       holds under NEXT_STEP autoreset, where no final observation exists to bootstrap from
template <int BOOT>
__global__ void __launch_bounds__(kScanThreads)
gae_kernel(const float* __restrict__ values, const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
           const uint8_t* __restrict__ timeouts, const float* __restrict__ last_values, const float* __restrict__ boot,
           int T, int64_t N, float gamma, float gl, float* __restrict__ adv, float* __restrict__ ret) {
    const int64_t n = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
    if (n >= N) return;
    float g = 0.0f;
    float vnext = __ldg(last_values + n);
    int t = T - 1;
    // software-pipelined: the loads of the NEXT kUnroll timesteps are issued before the (sequential, ~50-cycle-per-step) recurrence over the
    // current ones, so a thread keeps 2 x kUnroll timesteps of loads in flight -- at C2 an SM holds only ~14 warps of this kernel
    float v[kUnroll], r[kUnroll], bt[kUnroll];
    uint8_t d[kUnroll], to[kUnroll];
    auto load = [&](int t0, float (&v_)[kUnroll], float (&r_)[kUnroll], float (&bt_)[kUnroll], uint8_t (&d_)[kUnroll], uint8_t (&to_)[kUnroll]) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t o = (int64_t)(t0 - u) * N + n;
            v_[u] = ldg_stream(values + o);
            r_[u] = ldg_stream(rewards + o);
            d_[u] = ldg_stream(dones + o);
            to_[u] = ldg_stream(timeouts + o);
            bt_[u] = BOOT == 1 ? ldg_stream(boot + o) : 0.0f;
        }
    };
    if (t >= kUnroll - 1) load(t, v, r, bt, d, to);
    for (; t >= kUnroll - 1; t -= kUnroll) {
        float v2[kUnroll], r2[kUnroll], bt2[kUnroll];
        uint8_t d2[kUnroll], to2[kUnroll];
        const bool more = t - kUnroll >= kUnroll - 1;
        if (more) load(t - kUnroll, v2, r2, bt2, d2, to2);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t o = (int64_t)(t - u) * N + n;
            const float nv = (BOOT != 0 && to[u]) ? bt[u] : vnext;
            const float nt = (d[u] && !to[u]) ? 0.0f : 1.0f;
            const float delta = fsub(fadd(r[u], fmul(fmul(gamma, nv), nt)), v[u]);
            g = fadd(delta, fmul(fmul(gl, g), nt));
            stg_stream(adv + o, g);
            stg_stream(ret + o, fadd(g, v[u]));
            vnext = v[u];
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) { v[u] = v2[u]; r[u] = r2[u]; bt[u] = bt2[u]; d[u] = d2[u]; to[u] = to2[u]; }
        }
    }
    for (; t >= 0; --t) {
        const int64_t o = (int64_t)t * N + n;
        const float v = ldg_stream(values + o), r = ldg_stream(rewards + o);
        const uint8_t d = ldg_stream(dones + o), to = ldg_stream(timeouts + o);
        const float nv = (BOOT != 0 && to) ? (BOOT == 1 ? ldg_stream(boot + o) : 0.0f) : vnext;
        const float nt = (d && !to) ? 0.0f : 1.0f;
        const float delta = fsub(fadd(r, fmul(fmul(gamma, nv), nt)), v);
        g = fadd(delta, fmul(fmul(gl, g), nt));
        stg_stream(adv + o, g);
        stg_stream(ret + o, fadd(g, v));
        vnext = v;
    }
}

// reward-to-go (+ optional per-episode constant) and the last real terminal of every env
template <bool HAS_TO>
__global__ void __launch_bounds__(kScanThreads)
mc_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones, const uint8_t* __restrict__ timeouts, int T,
          int64_t N, float gamma, int episode_mode, float* __restrict__ ret, int32_t* __restrict__ last_terminal) {
    const int64_t n = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
    if (n >= N) return;
    float acc = 0.0f;
    int32_t last = -1;
    int t = T - 1;
    // software-pipelined like gae_kernel: the next batch's loads are in flight during this batch's recurrence
    float r[kUnroll];
    uint8_t d[kUnroll], to[kUnroll];
    auto load = [&](int t0, float (&r_)[kUnroll], uint8_t (&d_)[kUnroll], uint8_t (&to_)[kUnroll]) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t o = (int64_t)(t0 - u) * N + n;
            r_[u] = ldg_stream(rewards + o);
            d_[u] = ldg_stream(dones + o);
            to_[u] = HAS_TO ? ldg_stream(timeouts + o) : (uint8_t)0;
        }
    };
    if (t >= kUnroll - 1) load(t, r, d, to);
    for (; t >= kUnroll - 1; t -= kUnroll) {
        float r2[kUnroll];
        uint8_t d2[kUnroll], to2[kUnroll];
        const bool more = t - kUnroll >= kUnroll - 1;
        if (more) load(t - kUnroll, r2, d2, to2);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const bool term = d[u] && !to[u];
            if (term && last < 0) last = t - u;
            acc = fadd(r[u], fmul(gamma, fmul(acc, term ? 0.0f : 1.0f)));
            ret[(int64_t)(t - u) * N + n] = acc;
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) { r[u] = r2[u]; d[u] = d2[u]; to[u] = to2[u]; }
        }
    }
    for (; t >= 0; --t) {
        const int64_t o = (int64_t)t * N + n;
        const bool term = ldg_stream(dones + o) && !(HAS_TO && ldg_stream(timeouts + o));
        if (term && last < 0) last = t;
        acc = fadd(ldg_stream(rewards + o), fmul(gamma, fmul(acc, term ? 0.0f : 1.0f)));
        ret[o] = acc;
    }
    if (last_terminal) last_terminal[n] = last;
    if (episode_mode) {
        // every step of a segment takes the reward-to-go of the segment's first step; a segment ends at a real terminal
        float seg = 0.0f;
        bool start = true;
        for (int tt = 0; tt < T; ++tt) {
            const int64_t o = (int64_t)tt * N + n;
            if (start) seg = ret[o];
            else ret[o] = seg;
            start = __ldg(dones + o) && !(HAS_TO && __ldg(timeouts + o));
        }
    }
}

// convert_returns_to_full_episode as a standalone in-place pass (the fused path is mc_kernel's episode_mode)
template <bool HAS_TO>
__global__ void __launch_bounds__(kScanThreads)
full_episode_kernel(float* __restrict__ ret, const uint8_t* __restrict__ dones, const uint8_t* __restrict__ timeouts, int T, int64_t N) {
    const int64_t n = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
    if (n >= N) return;
    float seg = 0.0f;
    bool start = true;
    for (int t = 0; t < T; ++t) {
        const int64_t o = (int64_t)t * N + n;
        if (start) seg = ret[o];
        else ret[o] = seg;
        start = __ldg(dones + o) && !(HAS_TO && __ldg(timeouts + o));
    }
}

// ---- valid mask / index map ------------------------------------------------------------------------------------------
// inc[e] = largest env index e' <= e that has a real terminal (-1 if none): block-local inclusive max-scan + block maxima
__global__ void valid_scan_local_kernel(const int32_t* __restrict__ last_terminal, int64_t N, int32_t* __restrict__ inc,
                                        int32_t* __restrict__ block_max, unsigned long long* __restrict__ n_valid,
                                        int32_t* __restrict__ first_env) {
    __shared__ int32_t warp_max_s[32];
    __shared__ unsigned long long cnt_s[32];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t lt = e < N ? last_terminal[e] : -1;
    int32_t v = lt >= 0 ? (int32_t)e : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, u);
    }
    if (lane == 31) warp_max_s[w] = v;
    unsigned long long c = warp_sum((unsigned long long)(lt + 1));
    if (lane == 0) cnt_s[w] = c;
    __syncthreads();
    int32_t prefix = -1;
    for (int i = 0; i < w; ++i) prefix = max(prefix, warp_max_s[i]);
    v = max(v, prefix);
    if (e < N) inc[e] = v;
    if (threadIdx.x == blockDim.x - 1) block_max[blockIdx.x] = v;
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += cnt_s[i];
        if (tot) atomicAdd(n_valid, tot);
    }
    if (lt >= 0) atomicMin(first_env, (int32_t)e);
}

__global__ void valid_scan_blocks_kernel(int32_t* __restrict__ block_max, int n_blocks) {
    // exclusive max-scan over block maxima, in place (n_blocks <= N/1024: a few thousand entries)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int32_t run = -1;
    for (int b = 0; b < n_blocks; ++b) {
        const int32_t m = block_max[b];
        block_max[b] = run;
        run = max(run, m);
    }
}

__global__ void valid_apply_kernel(const int32_t* __restrict__ last_terminal, const int32_t* __restrict__ inc,
                                   const int32_t* __restrict__ block_prefix, const int32_t* __restrict__ first_env, int T,
                                   int64_t N, uint8_t* __restrict__ valid_mask, int64_t* __restrict__ idx_map) {
    const int64_t total = N * (int64_t)T;
    const int32_t fe = *first_env;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / T;
        const int t = (int)(i - e * T);
        const int32_t lt = last_terminal[e];
        int64_t target;
        uint8_t valid = 0;
        if (lt >= 0) {
            valid = t <= lt;
            target = valid ? i : e * T + lt;
        } else {
            const int32_t pe = max(inc[e], block_prefix[e >> 10]);
            if (pe >= 0) target = (int64_t)pe * T + last_terminal[pe];
            else target = fe < 0x7fffffff ? (int64_t)fe * T : i;   // prefix before the first valid entry -> first valid; NO valid entry at
                                                                   // all -> identity (the reference returns (None, None) and does not remap)
        }
        if (valid_mask) valid_mask[i] = valid;
        idx_map[i] = target;
    }
}
__global__ void valid_init_kernel(unsigned long long* n_valid, int32_t* first_env) { *n_valid = 0ull; *first_env = 0x7fffffff; }

// ---- moments / normalise ------------------------------------------------------------------------------------------------
__global__ void moments_kernel(const float* __restrict__ x, const int32_t* __restrict__ last_terminal, int T, int64_t N,
                               double* __restrict__ out, const int64_t* __restrict__ n_valid) {
    __shared__ double scratch[32];
    if (n_valid && *n_valid == 0) last_terminal = nullptr;   // no valid entry: the reference's mask is None and the statistics cover every element
    double s = 0.0, s2 = 0.0, c = 0.0;
    const int64_t total = N * (int64_t)T;
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (!last_terminal && (((uintptr_t)x) & 15) == 0) {
        // unmasked: 16-byte streaming loads, two in flight per thread (a 4-byte load per iteration left the kernel latency-bound
        // at ~1.2 TB/s); fp64 accumulation as before
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int64_t n4 = total >> 2;
        int64_t i = tid0;
        for (; i + stride < n4; i += 2 * stride) {
            const float4 a = __ldcs(x4 + i), b = __ldcs(x4 + i + stride);
            const double a0 = a.x, a1 = a.y, a2 = a.z, a3 = a.w, b0 = b.x, b1 = b.y, b2 = b.z, b3 = b.w;
            s += ((a0 + a1) + (a2 + a3)) + ((b0 + b1) + (b2 + b3));
            s2 += ((a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3)) + ((b0 * b0 + b1 * b1) + (b2 * b2 + b3 * b3));
            c += 8.0;
        }
        for (; i < n4; i += stride) {
            const float4 a = __ldcs(x4 + i);
            const double a0 = a.x, a1 = a.y, a2 = a.z, a3 = a.w;
            s += (a0 + a1) + (a2 + a3);
            s2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            c += 4.0;
        }
        for (int64_t j = (n4 << 2) + tid0; j < total; j += stride) {
            const double v = (double)ldg_stream(x + j);
            s += v; s2 += v * v; c += 1.0;
        }
    } else {
        for (int64_t i = tid0; i < total; i += stride) {
            bool use = true;
            if (last_terminal) {
                const int64_t t = i / N, n = i - t * N;
                use = t <= last_terminal[n];
            }
            if (use) {
                const double v = (double)ldg_stream(x + i);
                s += v; s2 += v * v; c += 1.0;
            }
        }
    }
    s = block_sum(s, scratch);
    s2 = block_sum(s2, scratch);
    c = block_sum(c, scratch);
    if (threadIdx.x == 0 && c > 0.0) { atomicAdd(out, s); atomicAdd(out + 1, s2); atomicAdd(out + 2, c); }
}

// x == y (in place) is allowed and used by the collector: no __restrict__ on the two
__global__ void normalize_kernel(const float* x, int64_t n, const double* __restrict__ mom, float eps, int mode, float* y) {
    const double cnt = mom[2] > 0.0 ? mom[2] : 1.0;
    const double mu = mom[0] / cnt;
    double var = mom[1] / cnt - mu * mu;   // population variance (numpy .std())
    var = var > 0.0 ? var : 0.0;
    const float mean = (float)mu;
    const float denom = mode == 0 ? (float)sqrt(var) + eps : 1.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = mode == 0 ? (x[i] - mean) / denom : x[i] - mean;
}

}  // namespace gs

using namespace gs;

extern "C" {

int gs_gae(const float* values, const float* rewards, const uint8_t* dones, const uint8_t* timeouts, const float* last_values,
           const float* bootstrapped, int T, int64_t N, double gamma, double gae_lambda, float* adv, float* ret, void* stream) {
    if (!values || !rewards || !dones || !timeouts || !last_values || !adv || !ret) GS_FAIL("gs_gae: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_gae: empty rollout (T=%d, N=%lld)", T, (long long)N);
    const unsigned blocks = (unsigned)((N + kScanThreads - 1) / kScanThreads);
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    cudaStream_t st = (cudaStream_t)stream;
    if (bootstrapped) gae_kernel<1><<<blocks, kScanThreads, 0, st>>>(values, rewards, dones, timeouts, last_values, bootstrapped, T, N, g, gl, adv, ret);
    else gae_kernel<0><<<blocks, kScanThreads, 0, st>>>(values, rewards, dones, timeouts, last_values, nullptr, T, N, g, gl, adv, ret);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_gae_zero_boot(const float* values, const float* rewards, const uint8_t* dones, const uint8_t* timeouts, const float* last_values,
                     int T, int64_t N, double gamma, double gae_lambda, float* adv, float* ret, void* stream) {
    if (!values || !rewards || !dones || !timeouts || !last_values || !adv || !ret) GS_FAIL("gs_gae_zero_boot: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_gae_zero_boot: empty rollout (T=%d, N=%lld)", T, (long long)N);
    const unsigned blocks = (unsigned)((N + kScanThreads - 1) / kScanThreads);
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    gae_kernel<2><<<blocks, kScanThreads, 0, (cudaStream_t)stream>>>(values, rewards, dones, timeouts, last_values, nullptr, T, N, g, gl, adv, ret);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_mc_returns(const float* rewards, const uint8_t* dones, const uint8_t* timeouts, int T, int64_t N, double gamma,
                  int episode_mode, float* ret, int32_t* last_terminal, void* stream) {
    if (!rewards || !dones || !ret) GS_FAIL("gs_mc_returns: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_mc_returns: empty rollout (T=%d, N=%lld)", T, (long long)N);
    const unsigned blocks = (unsigned)((N + kScanThreads - 1) / kScanThreads);
    cudaStream_t st = (cudaStream_t)stream;
    if (timeouts) mc_kernel<true><<<blocks, kScanThreads, 0, st>>>(rewards, dones, timeouts, T, N, (float)gamma, episode_mode, ret, last_terminal);
    else mc_kernel<false><<<blocks, kScanThreads, 0, st>>>(rewards, dones, nullptr, T, N, (float)gamma, episode_mode, ret, last_terminal);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_returns_to_full_episode(float* ret, const uint8_t* dones, const uint8_t* timeouts, int T, int64_t N, void* stream) {
    if (!ret || !dones) GS_FAIL("gs_returns_to_full_episode: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_returns_to_full_episode: empty rollout");
    const unsigned blocks = (unsigned)((N + kScanThreads - 1) / kScanThreads);
    cudaStream_t st = (cudaStream_t)stream;
    if (timeouts) full_episode_kernel<true><<<blocks, kScanThreads, 0, st>>>(ret, dones, timeouts, T, N);
    else full_episode_kernel<false><<<blocks, kScanThreads, 0, st>>>(ret, dones, nullptr, T, N);
    GS_LAUNCH_CHECK();
    return 0;
}

int64_t gs_valid_index_map_workspace_bytes(int64_t N) {
    const int64_t nb = (N + 1023) / 1024;
    return ((N * 4 + 255) / 256 + (nb * 4 + 255) / 256 + 1) * 256;
}

int gs_valid_index_map(const int32_t* last_terminal, int T, int64_t N, uint8_t* valid_mask, int64_t* idx_map, int64_t* n_valid,
                       void* workspace, int64_t workspace_bytes, void* stream) {
    if (!last_terminal || !idx_map || !n_valid || !workspace) GS_FAIL("gs_valid_index_map: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_valid_index_map: empty rollout");
    if (N >= (1ll << 31)) GS_FAIL("gs_valid_index_map: N too large");
    if (workspace_bytes < gs_valid_index_map_workspace_bytes(N)) GS_FAIL("gs_valid_index_map: workspace too small");
    const int64_t nb = (N + 1023) / 1024;
    char* p = (char*)workspace;
    int32_t* inc = (int32_t*)p; p += (N * 4 + 255) / 256 * 256;
    int32_t* block_max = (int32_t*)p; p += (nb * 4 + 255) / 256 * 256;
    int32_t* first_env = (int32_t*)p;
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    valid_init_kernel<<<1, 1, 0, st>>>((unsigned long long*)n_valid, first_env);
    valid_scan_local_kernel<<<(unsigned)nb, 1024, 0, st>>>(last_terminal, N, inc, block_max, (unsigned long long*)n_valid, first_env);
    valid_scan_blocks_kernel<<<1, 32, 0, st>>>(block_max, (int)nb);
    const int64_t total = N * (int64_t)T;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = 8ll * sm_count(device);
    if (blocks > cap) blocks = cap;
    valid_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(last_terminal, inc, block_max, first_env, T, N, valid_mask, idx_map);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_moments(const float* x, const int32_t* last_terminal, int T, int64_t N, double* out, void* stream) {
    if (!x || !out) GS_FAIL("gs_moments: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_moments: empty input");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    const int64_t total = N * (int64_t)T;
    int64_t blocks = (total + 1023) / 1024;
    const int64_t cap = 8ll * sm_count(device);
    if (blocks > cap) blocks = cap;
    moments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, last_terminal, T, N, out, nullptr);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_moments_valid(const float* x, const int32_t* last_terminal, const int64_t* n_valid, int T, int64_t N, double* out, void* stream) {
    if (!x || !out || !last_terminal || !n_valid) GS_FAIL("gs_moments_valid: NULL argument");
    if (T <= 0 || N <= 0) GS_FAIL("gs_moments_valid: empty input");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    const int64_t total = N * (int64_t)T;
    int64_t blocks = (total + 1023) / 1024;
    const int64_t cap = 8ll * sm_count(device);
    if (blocks > cap) blocks = cap;
    moments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, last_terminal, T, N, out, n_valid);
    GS_LAUNCH_CHECK();
    return 0;
}

static int launch_normalize(const float* x, int64_t n, const double* moments, float eps, int mode, float* y, void* stream) {
    if (!x || !y || !moments) GS_FAIL("normalize: NULL argument");
    if (n <= 0) GS_FAIL("normalize: empty input");
    int device = 0;
    GS_CUDA(cudaGetDevice(&device));
    int64_t blocks = (n + 1023) / 1024;
    const int64_t cap = 8ll * sm_count(device);
    if (blocks > cap) blocks = cap;
    normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, moments, eps, mode, y);
    GS_LAUNCH_CHECK();
    return 0;
}

int gs_normalize(const float* x, int64_t n, const double* moments, float eps, float* y, void* stream) {
    return launch_normalize(x, n, moments, eps, 0, y, stream);
}
int gs_shift_by_mean(const float* x, int64_t n, const double* moments, float* y, void* stream) {
    return launch_normalize(x, n, moments, 0.f, 1, y, stream);
}

}  // extern "C"
