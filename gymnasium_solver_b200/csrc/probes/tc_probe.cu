// tc_probe.cu — hardware probe for the tcgen05 conventions the tensor-core update kernel relies on (dev tool, not shipped
// in libgs_engine.so).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run on a B200.
//   T1  K-major A (M=128) x K-major B (N=64), K=64, 3xTF32 split         -> forward GEMM
//   T2  K-major A (M=128) x MN-major B (N=64), K=64                       -> dgrad with W in its natural layout
//   T3  MN-major A (M=64) x MN-major B (N=72), K=128, accumulated twice   -> wgrad (+ bias column); dumps the M=64 TMEM lanes
//   T4  MN-major A (M=64) x no-swizzle MN-major B (N=8), K=128            -> small-N grads
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"

using namespace gs::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

// smem carve (floats): A_hi[128x64] A_lo, B_hi[128x96] B_lo (generous)
constexpr int kAF = 128 * 64, kBF = 128 * 96;

__device__ void store_split(float* hi, float* lo, int idx, float v) {
    float h, l; split_tf32(v, h, l); hi[idx] = h; lo[idx] = l;
}

// mode 1: T1, 2: T2, 3: T3, 4: T4
__global__ void probe_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D /* [128 lanes][128 cols] dump */) {
    extern __shared__ __align__(1024) float sm[];
    float* a_hi = sm; float* a_lo = a_hi + kAF; float* b_hi = a_lo + kAF; float* b_lo = b_hi + kBF;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    // ---- stage operands -------------------------------------------------------------------------------------------
    if (mode == 1) {           // A[128][64] rows=M ; B[64 n][64 k] rows=N
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(a_hi, a_lo, slab_index(r, c, 128), A[i]); }
        for (int i = tid; i < 64 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(b_hi, b_lo, slab_index(r, c, 64), B[i]); }
    } else if (mode == 2) {    // A = G[128 s][64 n] ; B = W[64 n][64 k] natural
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(a_hi, a_lo, slab_index(r, c, 128), A[i]); }
        for (int i = tid; i < 64 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(b_hi, b_lo, slab_index(r, c, 64), B[i]); }
    } else if (mode == 3) {    // A = G[128 s][64 n] ; B = X[128 s][72]
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(a_hi, a_lo, slab_index(r, c, 128), A[i]); }
        for (int i = tid; i < 128 * 96; i += 128) { b_hi[i] = 0.f; b_lo[i] = 0.f; }
        __syncthreads();
        for (int i = tid; i < 128 * 72; i += 128) { int r = i / 72, c = i % 72; store_split(b_hi, b_lo, slab_index(r, c, 128), B[i]); }
    } else if (mode == 5) {    // TS: A[128][64] -> TMEM cols 0..63 (hi) / 64..127 (lo) by tcgen05.st ; B[64][64] K-major smem
        for (int i = tid; i < 64 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(b_hi, b_lo, slab_index(r, c, 64), B[i]); }
    } else if (mode == 6 || mode == 7) {   // K-major transposed wgrad: A = GT[64 n][128 s], B = XT[N rows][128 s], N = 72 (mode 6) or 8 (mode 7)
        const int N = mode == 6 ? 72 : 8;
        for (int i = tid; i < 128 * 64; i += 128) { int s_ = i / 64, n = i % 64; store_split(a_hi, a_lo, slab_index(n, s_, 64), A[i]); }
        for (int i = tid; i < 128 * N; i += 128) { int s_ = i / N, j = i % N; store_split(b_hi, b_lo, slab_index(j, s_, N), B[i]); }
    } else if (mode == 8) {    // T1 again but the hi operand holds the FULL fp32 value (does the tensor core truncate or round?)
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; float h, l; split_tf32(A[i], h, l); a_hi[slab_index(r, c, 128)] = A[i]; a_lo[slab_index(r, c, 128)] = l; }
        for (int i = tid; i < 64 * 64; i += 128) { int r = i / 64, c = i % 64; float h, l; split_tf32(B[i], h, l); b_hi[slab_index(r, c, 64)] = B[i]; b_lo[slab_index(r, c, 64)] = l; }
    } else {                   // A = G[128 s][64 n] ; B = X8[128 s][8] no swizzle: [s/8][chunk c][s%8][4]
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; store_split(a_hi, a_lo, slab_index(r, c, 128), A[i]); }
        for (int i = tid; i < 128 * 8; i += 128) { int r = i / 8, c = i % 8; store_split(b_hi, b_lo, (r / 8) * 64 + (c / 4) * 32 + (r % 8) * 4 + (c % 4), B[i]); }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (mode == 5) {           // every thread writes its own row (lane) of A into TMEM
        for (int c0 = 0; c0 < 64; c0 += 16) {
            float h[16], l[16];
            for (int i = 0; i < 16; ++i) split_tf32(A[tid * 64 + c0 + i], h[i], l[i]);
            tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c0, h);
            tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c0, l);
        }
        tmem_st_wait();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    const uint32_t dcol = mode == 5 ? 128u : 0u;
    // ---- issue ---------------------------------------------------------------------------------------------------------
    if (tid == 0) {
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
        if (mode == 1) {
            const uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
            int first = 1;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a0 = pass == 1 ? al : ah, b0 = pass == 2 ? bl : bh;
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t aoff = (kk >> 2) * 128 * 128 + (kk & 3) * 32, boff = (kk >> 2) * 64 * 128 + (kk & 3) * 32;
                    mma_tf32(tmem, make_desc(a0 + aoff, 16, 1024, kLayoutSW128), make_desc(b0 + boff, 16, 1024, kLayoutSW128), idesc, first ? 0 : 1);
                    first = 0;
                }
            }
        } else if (mode == 2) {
            const uint32_t idesc = make_idesc_tf32(128, 64, 0, 1);
            int first = 1;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a0 = pass == 1 ? al : ah, b0 = pass == 2 ? bl : bh;
                for (int kk = 0; kk < 8; ++kk) {   // K = n: 8 rows of W per MMA
                    const uint32_t aoff = (kk >> 2) * 128 * 128 + (kk & 3) * 32, boff = kk * 1024;
                    mma_tf32(tmem, make_desc(a0 + aoff, 16, 1024, kLayoutSW128), make_desc(b0 + boff, 64 * 128, 1024, kLayoutSW128), idesc, first ? 0 : 1);
                    first = 0;
                }
            }
        } else if (mode == 3) {
            const uint32_t idesc = make_idesc_tf32(64, 72, 1, 1);
            int first = 1;
            for (int rep = 0; rep < 2; ++rep)
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a0 = pass == 1 ? al : ah, b0 = pass == 2 ? bl : bh;
                    for (int kk = 0; kk < 16; ++kk) {   // K = s: 8 sample rows per MMA
                        mma_tf32(tmem, make_desc(a0 + kk * 1024, 128 * 128, 1024, kLayoutSW128), make_desc(b0 + kk * 1024, 128 * 128, 1024, kLayoutSW128), idesc, first ? 0 : 1);
                        first = 0;
                    }
                }
        } else if (mode == 5) {
            const uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
            int first = 1;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t acol = pass == 1 ? 64u : 0u, b0 = pass == 2 ? bl : bh;
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t boff = (kk >> 2) * 64 * 128 + (kk & 3) * 32;
                    mma_tf32_ts(tmem + dcol, tmem + acol + kk * 8, make_desc(b0 + boff, 16, 1024, kLayoutSW128), idesc, first ? 0 : 1);
                    first = 0;
                }
            }
        } else if (mode == 6 || mode == 7) {
            const int N = mode == 6 ? 72 : 8;
            const uint32_t idesc = make_idesc_tf32(64, N, 0, 0);
            int first = 1;
            for (int rep = 0; rep < 2; ++rep)
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a0 = pass == 1 ? al : ah, b0 = pass == 2 ? bl : bh;
                    for (int kk = 0; kk < 16; ++kk) {   // K = s: 8 samples per MMA; 4 slabs of 32 samples
                        const uint32_t aoff = (kk >> 2) * 64 * 128 + (kk & 3) * 32, boff = (kk >> 2) * N * 128 + (kk & 3) * 32;
                        mma_tf32(tmem, make_desc(a0 + aoff, 16, 1024, kLayoutSW128), make_desc(b0 + boff, 16, 1024, kLayoutSW128), idesc, first ? 0 : 1);
                        first = 0;
                    }
                }
        } else if (mode == 8) {
            const uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
            int first = 1;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a0 = pass == 1 ? al : ah, b0 = pass == 2 ? bl : bh;
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t aoff = (kk >> 2) * 128 * 128 + (kk & 3) * 32, boff = (kk >> 2) * 64 * 128 + (kk & 3) * 32;
                    mma_tf32(tmem, make_desc(a0 + aoff, 16, 1024, kLayoutSW128), make_desc(b0 + boff, 16, 1024, kLayoutSW128), idesc, first ? 0 : 1);
                    first = 0;
                }
            }
        } else {
            const uint32_t idesc = make_idesc_tf32(64, 8, 1, 1);
            int first = 1;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a0 = pass == 1 ? al : ah, b0 = pass == 2 ? bl : bh;
                for (int kk = 0; kk < 16; ++kk) {
                    // no-swizzle MN-major: LBO = k-group stride (256 B), SBO = MN-chunk stride (128 B)
                    mma_tf32(tmem, make_desc(a0 + kk * 1024, 128 * 128, 1024, kLayoutSW128), make_desc(b0 + kk * 256, 256, 128, kLayoutNone), idesc, first ? 0 : 1);
                    first = 0;
                }
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    // ---- dump all 128 lanes x 96 columns ---------------------------------------------------------------------------------
    for (int c0 = 0; c0 < 96; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + dcol + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) D[tid * 128 + c0 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
    const size_t smem = (size_t)(2 * kAF + 2 * kBF) * 4 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> A(128 * 64), B(128 * 96), D(128 * 128);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    int fails = 0;
    const int modes[] = {1, 8, 5, 6, 7};
    for (int mi = 0; mi < 5; ++mi) {
        const int mode = modes[mi];
        srand(mode);
        for (auto& x : A) x = frand();
        for (auto& x : B) x = frand();
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0, D.size() * 4));
        probe_kernel<<<1, 128, smem>>>(mode, dA, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T%d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0;
        if (mode == 1 || mode == 2 || mode == 5 || mode == 8) {
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 64; ++k) ref += mode != 2 ? (double)A[m * 64 + k] * B[n * 64 + k] : (double)A[m * 64 + k] * B[k * 64 + n];
                    max_err = fmax(max_err, fabs(ref - D[m * 128 + n])); max_ref = fmax(max_ref, fabs(ref));
                }
        } else {
            const int N = (mode == 3 || mode == 6) ? 72 : 8;
            const double rep = (mode == 3 || mode == 6 || mode == 7) ? 2.0 : 1.0;
            // discover which TMEM lane holds output row m: try lane = m (dense), lane = (m/16)*32 + m%16 (16 per quadrant), lane = (m/32)*... report best
            for (int layout = 0; layout < 3; ++layout) {
                double err = 0, mref = 0;
                for (int m = 0; m < 64; ++m) {
                    const int lane = layout == 0 ? m : (layout == 1 ? (m / 16) * 32 + (m % 16) : (m / 32) * 64 + (m % 32));
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int s = 0; s < 128; ++s) ref += (double)A[s * 64 + m] * B[s * N + n];
                        ref *= rep;
                        err = fmax(err, fabs(ref - D[lane * 128 + n])); mref = fmax(mref, fabs(ref));
                    }
                }
                printf("T%d: M=64 lane layout hypothesis %d: max_err %.3e (max_ref %.3e)\n", mode, layout, err, mref);
                if (layout == 0 || err < max_err) { max_err = err; max_ref = mref; }
            }
            int used = 0;
            for (int l = 0; l < 128; ++l) { bool nz = false; for (int n = 0; n < N; ++n) nz |= D[l * 128 + n] != 0.f; used += nz; if (nz && (l % 16 == 0)) printf("    lane %d non-zero\n", l); }
            printf("T%d: %d TMEM lanes hold data\n", mode, used);
        }
        const bool ok = max_err <= 2e-5 * fmax(1.0, max_ref);
        printf("T%d: max_err %.3e max_ref %.3e -> %s\n", mode, max_err, max_ref, ok ? "OK" : "MISMATCH");
        fails += !ok;
    }
    printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
    return fails ? 1 : 0;
}
