// bf16_probe.cu — known-answer probe of the tcgen05 kind::f16 (bf16) operand conventions the bf16x3 update kernel relies on
// (dev tool, not part of libgs_engine.so).  The host builds a shared-memory IMAGE and a list of MMA instructions (descriptor
// offsets relative to the image), the kernel copies the image, issues the list from one thread and dumps TMEM; every layout
// decision therefore lives in host code and one kernel serves all cases.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o bf16_probe bf16_probe.cu ; run on a B200.
// Tile convention under test (the ONE physical layout of every activation tile): [rows][64*S] bf16, S slabs of [rows][64]
// (128-byte rows), 16-byte chunk c of row r stored at chunk (c ^ (r & 7)):
//     byte(r, c) = (c / 64) * rows * 128 + r * 128 + ((((c % 64) / 8) ^ (r & 7)) * 16) + (c % 8) * 2
//   * K-major SWIZZLE_128B operand with MN = r, K = c   (SBO = 1024; k-step of 16 = +32 bytes, next slab after 4 k-steps)
//   * MN-major SWIZZLE_128B operand with MN = c, K = r  (SBO = 1024 between 8-row groups, LBO = slab stride)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../tc_common.cuh"

using namespace gs::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct Mma { uint32_t a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, idesc, dcol, acc; };

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int kImage = 160 * 1024;

__global__ void probe_kernel(const unsigned char* __restrict__ image, int image_bytes, const Mma* __restrict__ list, int n_mma,
                             float* __restrict__ dump /* [128 lanes][256 cols] */) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < image_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(image)[i];
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    // zero the TMEM dump region first so untouched lanes / columns read as 0
    {
        float z[16];
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
        for (int c = 0; c < 256; c += 16) tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c, z);
        tmem_st_wait();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    if (tid == 0) {
        const uint32_t base = smem_u32(sm);
        for (int i = 0; i < n_mma; ++i) {
            const Mma m = list[i];
            mma_f16(tmem + m.dcol, make_desc(base + m.a_off, m.a_lbo, m.a_sbo, kLayoutSW128), make_desc(base + m.b_off, m.b_lbo, m.b_sbo, kLayoutSW128),
                    m.idesc, m.acc);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c = 0; c < 256; c += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) dump[tid * 256 + c + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static uint16_t bf16_bits(float x) { __nv_bfloat16 b = __float2bfloat16(x); uint16_t u; memcpy(&u, &b, 2); return u; }
static size_t tile_byte(int r, int c, int rows) { return (size_t)(c / 64) * rows * 128 + (size_t)r * 128 + (size_t)((((c % 64) / 8) ^ (r & 7)) * 16) + (c % 8) * 2; }

static float f16_round(float x) { return __half2float(__float2half(x)); }
static uint16_t f16_bits(float x) { __half b = __float2half(x); uint16_t u; memcpy(&u, &b, 2); return u; }
struct Tile {   // [rows][cols] bf16 (or fp16) values (already rounded) placed at image offset `off`
    int rows, cols; uint32_t off; std::vector<float> v; bool half = false;
    void fill16(unsigned seed) { half = true; srand(seed); for (auto& x : v) x = f16_round(((rand() % 200001) - 100000) / 100000.0f); }
    Tile(int r, int c, uint32_t o) : rows(r), cols(c), off(o), v((size_t)r * c, 0.f) {}
    float& at(int r, int c) { return v[(size_t)r * cols + c]; }
    void fill(unsigned seed, float scale = 1.f) { srand(seed); for (auto& x : v) x = bf16_round(scale * ((rand() % 2001) - 1000) / 1000.0f); }
    void put(std::vector<unsigned char>& img) {
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) { uint16_t b = half ? f16_bits(at(r, c)) : bf16_bits(at(r, c)); memcpy(&img[off + tile_byte(r, c, rows)], &b, 2); }
    }
    uint32_t slab() const { return (uint32_t)rows * 128; }
};

static std::vector<unsigned char> g_img;
static std::vector<Mma> g_list;
static float* d_dump; static unsigned char* d_img; static Mma* d_list;
static std::vector<float> h_dump(128 * 256);

static void run() {
    CK(cudaMemcpy(d_img, g_img.data(), kImage, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_list, g_list.data(), g_list.size() * sizeof(Mma), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_dump, 0, 128 * 256 * 4));
    probe_kernel<<<1, 128, kImage>>>(d_img, kImage, d_list, (int)g_list.size(), d_dump);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_dump.data(), d_dump, 128 * 256 * 4, cudaMemcpyDeviceToHost));
}
// lane of accumulator row m
static int lane_of(int m, int M) { return M == 128 ? m : (m / 16) * 32 + (m % 16); }
static bool report(const char* name, int M, int N, int dcol, const std::vector<double>& ref /* [M][N] */) {
    double worst = 0, scale = 0; int wm = 0, wn = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        const double got = h_dump[lane_of(m, M) * 256 + dcol + n], want = ref[(size_t)m * N + n];
        scale = fmax(scale, fabs(want));
        if (fabs(got - want) > worst) { worst = fabs(got - want); wm = m; wn = n; }
    }
    const bool ok = worst <= 2e-5 * fmax(scale, 1.0);
    printf("%-58s %s  max|err| %.3e (scale %.2f) at m=%d n=%d got %.6f want %.6f\n", name, ok ? "OK  " : "FAIL", worst, scale, wm, wn,
           h_dump[lane_of(wm, M) * 256 + dcol + wn], ref[(size_t)wm * N + wn]);
    return ok;
}
static void begin() { g_img.assign(kImage, 0); g_list.clear(); }

int main() {
    CK(cudaMalloc(&d_dump, 128 * 256 * 4)); CK(cudaMalloc(&d_img, kImage)); CK(cudaMalloc(&d_list, 4096 * sizeof(Mma)));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kImage));
    int fails = 0;

    // P1: forward.  z[s][j] = sum_k h[s][k] W[j][k].  A = h tile K-major (M=128, K=64), B = W tile [64 j][64 k] K-major.
    for (int H : {64, 128}) {
        begin();
        Tile h(128, H, 0), W(H, H, 65536);
        h.fill(1); W.fill(2); h.put(g_img); W.put(g_img);
        for (int kk = 0; kk < H / 16; ++kk)
            g_list.push_back({h.off + (kk / 4) * h.slab() + (kk % 4) * 32, 16, 1024, W.off + (kk / 4) * W.slab() + (kk % 4) * 32, 16, 1024,
                              idesc_bf16(128, H, 0, 0), 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)128 * H);
        for (int s = 0; s < 128; ++s) for (int j = 0; j < H; ++j) { double a = 0; for (int k = 0; k < H; ++k) a += (double)h.at(s, k) * W.at(j, k); ref[(size_t)s * H + j] = a; }
        char nm[96]; snprintf(nm, 96, "P1 fwd   K-major A x K-major B      H=%d", H);
        fails += !report(nm, 128, H, 0, ref);
    }
    // P2: dgrad.  dh[s][k] = sum_j dz[s][j] W[j][k].  A = dz tile K-major (K = j), B = the SAME W tile read MN-major (N = k, K = j).
    for (int H : {64, 128}) {
        begin();
        Tile dz(128, H, 0), W(H, H, 65536);
        dz.fill(3); W.fill(4); dz.put(g_img); W.put(g_img);
        for (int kk = 0; kk < H / 16; ++kk)   // K = j: 16 rows of W per MMA = 2 groups of 8 rows
            g_list.push_back({dz.off + (kk / 4) * dz.slab() + (kk % 4) * 32, 16, 1024, W.off + kk * 2048, W.slab(), 1024,
                              idesc_bf16(128, H, 0, 1), 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)128 * H);
        for (int s = 0; s < 128; ++s) for (int k = 0; k < H; ++k) { double a = 0; for (int j = 0; j < H; ++j) a += (double)dz.at(s, j) * W.at(j, k); ref[(size_t)s * H + k] = a; }
        char nm[96]; snprintf(nm, 96, "P2 dgrad K-major A x MN-major B     H=%d", H);
        fails += !report(nm, 128, H, 0, ref);
    }
    // P3: wgrad.  dW[j][k] = sum_s dz[s][j] h[s][k].  A = dz tile MN-major (M = j), B = h tile MN-major (N = k), K = 128 samples.
    for (int H : {64, 128}) {
        begin();
        Tile dz(128, H, 0), h(128, H, 65536);
        dz.fill(5); h.fill(6); dz.put(g_img); h.put(g_img);
        for (int kk = 0; kk < 8; ++kk)
            g_list.push_back({dz.off + kk * 2048, dz.slab(), 1024, h.off + kk * 2048, h.slab(), 1024, idesc_bf16(H, H, 1, 1), 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)H * H);
        for (int j = 0; j < H; ++j) for (int k = 0; k < H; ++k) { double a = 0; for (int s = 0; s < 128; ++s) a += (double)dz.at(s, j) * h.at(s, k); ref[(size_t)j * H + k] = a; }
        char nm[96]; snprintf(nm, 96, "P3 wgrad MN-major A x MN-major B    M=N=%d K=128", H);
        fails += !report(nm, H, H, 0, ref);
    }
    // X tile: [128 samples][64] bf16, 16-column groups g = 0..3 used as K=16 / N=16 operands at byte offset 32 g inside the row.
    // P5: layer 1.  z[s][j] = sum_c X[s][16g + c] WS[j][16q + c]: A = X K-major (K=16, start +32g), B = WS [H][64] K-major (start +32q).
    for (int H : {64, 128}) for (int g : {0, 2}) for (int q : {0, 1, 3}) {
        begin();
        Tile X(128, 64, 0), WS(H, 64, 65536);
        X.fill(7); WS.fill(8); X.put(g_img); WS.put(g_img);
        g_list.push_back({X.off + 32u * g, 16, 1024, WS.off + 32u * q, 16, 1024, idesc_bf16(128, H, 0, 0), 0, 0u});
        run();
        std::vector<double> ref((size_t)128 * H);
        for (int s = 0; s < 128; ++s) for (int j = 0; j < H; ++j) { double a = 0; for (int c = 0; c < 16; ++c) a += (double)X.at(s, 16 * g + c) * WS.at(j, 16 * q + c); ref[(size_t)s * H + j] = a; }
        char nm[96]; snprintf(nm, 96, "P5 K=16 sub-tile A(+%dB) x B(+%dB)   N=%d", 32 * g, 32 * q, H);
        fails += !report(nm, 128, H, 0, ref);
    }
    // P6: heads.  out[s][r] = sum_k h[s][k] T[r][k]: A = h K-major, B = T [16][H] K-major (N = 16).
    for (int H : {64, 128}) {
        begin();
        Tile h(128, H, 0), T(16, H, 65536);
        h.fill(9); T.fill(10); h.put(g_img); T.put(g_img);
        for (int kk = 0; kk < H / 16; ++kk)
            g_list.push_back({h.off + (kk / 4) * h.slab() + (kk % 4) * 32, 16, 1024, T.off + (kk / 4) * T.slab() + (kk % 4) * 32, 16, 1024,
                              idesc_bf16(128, 16, 0, 0), 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)128 * 16);
        for (int s = 0; s < 128; ++s) for (int r = 0; r < 16; ++r) { double a = 0; for (int k = 0; k < H; ++k) a += (double)h.at(s, k) * T.at(r, k); ref[(size_t)s * 16 + r] = a; }
        char nm[96]; snprintf(nm, 96, "P6 heads K-major A x K-major B N=16 H=%d", H);
        fails += !report(nm, 128, 16, 0, ref);
    }
    // P7: dh2[s][j] = sum_r G[s][r] T[r][j]: A = X group g (K = 16), B = T read MN-major (N = j, K = r: 16 rows).
    for (int H : {64, 128}) for (int g : {1, 3}) {
        begin();
        Tile X(128, 64, 0), T(16, H, 65536);
        X.fill(11); T.fill(12); X.put(g_img); T.put(g_img);
        g_list.push_back({X.off + 32u * g, 16, 1024, T.off, T.slab(), 1024, idesc_bf16(128, H, 0, 1), 0, 0u});
        run();
        std::vector<double> ref((size_t)128 * H);
        for (int s = 0; s < 128; ++s) for (int j = 0; j < H; ++j) { double a = 0; for (int r = 0; r < 16; ++r) a += (double)X.at(s, 16 * g + r) * T.at(r, j); ref[(size_t)s * H + j] = a; }
        char nm[96]; snprintf(nm, 96, "P7 dh2  A = X(+%dB) x MN-major T      N=%d", 32 * g, H);
        fails += !report(nm, 128, H, 0, ref);
    }
    // P8: small wgrad.  D[j][c] = sum_s dz[s][j] X[s][16g + c]: A = dz MN-major (M = H), B = X group g MN-major (N = 16, start +32g), K = 128.
    for (int H : {64, 128}) for (int g : {0, 1, 2}) {
        begin();
        Tile dz(128, H, 0), X(128, 64, 65536);
        dz.fill(13); X.fill(14); dz.put(g_img); X.put(g_img);
        for (int kk = 0; kk < 8; ++kk)
            g_list.push_back({dz.off + kk * 2048, dz.slab(), 1024, X.off + 32u * g + kk * 2048, X.slab(), 1024, idesc_bf16(H, 16, 1, 1), 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)H * 16);
        for (int j = 0; j < H; ++j) for (int c = 0; c < 16; ++c) { double a = 0; for (int s = 0; s < 128; ++s) a += (double)dz.at(s, j) * X.at(s, 16 * g + c); ref[(size_t)j * 16 + c] = a; }
        char nm[96]; snprintf(nm, 96, "P8 wgrad MN-major A x X(+%dB) N=16    M=%d", 32 * g, H);
        fails += !report(nm, H, 16, 0, ref);
    }
    // P9: merged wgrad + bias column for H=64: B = [h | X] (N = 80: second 64-wide atom found through LBO), A = dz (M = 64).
    {
        begin();
        Tile dz(128, 64, 0), h(128, 64, 16384), X(128, 64, 49152);
        dz.fill(15); h.fill(16); X.fill(17); dz.put(g_img); h.put(g_img); X.put(g_img);
        for (int kk = 0; kk < 8; ++kk)
            g_list.push_back({dz.off + kk * 2048, dz.slab(), 1024, h.off + kk * 2048, X.off - h.off, 1024, idesc_bf16(64, 80, 1, 1), 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)64 * 80);
        for (int j = 0; j < 64; ++j) for (int n = 0; n < 80; ++n) {
            double a = 0;
            for (int s = 0; s < 128; ++s) a += (double)dz.at(s, j) * (n < 64 ? h.at(s, n) : X.at(s, n - 64));
            ref[(size_t)j * 80 + n] = a;
        }
        fails += !report("P9 wgrad N=80 = [h | X] through LBO   M=64", 64, 80, 0, ref);
    }
    // P10: accumulate flag + a second accumulator at a 16-column offset (dcol = 80), M = 128 A = [h1 | h2] two different tiles via LBO
    {
        begin();
        Tile a0(128, 64, 0), a1(128, 64, 32768), X(128, 64, 65536);
        a0.fill(18); a1.fill(19); X.fill(20); a0.put(g_img); a1.put(g_img); X.put(g_img);
        for (int rep = 0; rep < 2; ++rep)
            for (int kk = 0; kk < 8; ++kk)
                g_list.push_back({a0.off + kk * 2048, a1.off - a0.off, 1024, X.off + 32 + kk * 2048, X.slab(), 1024, idesc_bf16(128, 16, 1, 1), 80, (rep | kk) ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)128 * 16);
        for (int m = 0; m < 128; ++m) for (int c = 0; c < 16; ++c) {
            double a = 0;
            for (int s = 0; s < 128; ++s) a += 2.0 * (double)(m < 64 ? a0.at(s, m) : a1.at(s, m - 64)) * X.at(s, 16 + c);
            ref[(size_t)m * 16 + c] = a;
        }
        fails += !report("P10 M=128 A=[t0 | t1] via LBO x X(+32B), 2x accumulate", 128, 16, 80, ref);
    }
    // PM: mixed operand formats in one kind::f16 MMA: (A, B) in {f16, bf16}^2 -- forward operands could be fp16 pairs (22 bits), backward bf16
    for (int fa : {0, 1}) for (int fb : {0, 1}) {
        begin();
        Tile h(128, 64, 0), W(64, 64, 65536);
        if (fa) h.fill(31); else h.fill16(31);
        if (fb) W.fill(32); else W.fill16(32);
        h.put(g_img); W.put(g_img);
        const uint32_t id = (1u << 4) | ((uint32_t)fa << 7) | ((uint32_t)fb << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kk = 0; kk < 4; ++kk) g_list.push_back({h.off + kk * 32, 16, 1024, W.off + kk * 32, 16, 1024, id, 0, kk > 0 ? 1u : 0u});
        run();
        std::vector<double> ref((size_t)128 * 64);
        for (int s = 0; s < 128; ++s) for (int j = 0; j < 64; ++j) { double a = 0; for (int k = 0; k < 64; ++k) a += (double)h.at(s, k) * W.at(j, k); ref[(size_t)s * 64 + j] = a; }
        char nm[96]; snprintf(nm, 96, "PM mixed formats A=%s x B=%s", fa ? "bf16" : "f16", fb ? "bf16" : "f16");
        fails += !report(nm, 128, 64, 0, ref);
    }
    // PB: accumulation bias.  Positive operands (no cancellation): mean signed relative error of a chain of n MMAs into one
    // accumulator against the exact (double) sum of the same bf16 products -> does the accumulator round or truncate, and by how much?
    for (int chain : {1, 2, 4, 8, 16, 32}) {
        begin();
        Tile h(128, 64, 0), W(64, 64, 65536);
        srand(100 + chain);
        for (auto& x : h.v) x = bf16_round(0.25f + (rand() % 1000) / 1000.0f);
        for (auto& x : W.v) x = bf16_round(0.25f + (rand() % 1000) / 1000.0f);
        h.put(g_img); W.put(g_img);
        for (int i = 0; i < chain; ++i) {
            const int kk = i % 4;
            g_list.push_back({h.off + kk * 32, 16, 1024, W.off + kk * 32, 16, 1024, idesc_bf16(128, 64, 0, 0), 0, i > 0 ? 1u : 0u});
        }
        run();
        double mean_rel = 0, rms = 0;
        for (int s = 0; s < 128; ++s) for (int j = 0; j < 64; ++j) {
            double a = 0;
            for (int i = 0; i < chain; ++i) { const int kk = i % 4; for (int k = 0; k < 16; ++k) a += (double)h.at(s, kk * 16 + k) * W.at(j, kk * 16 + k); }
            const double rel = (h_dump[s * 256 + j] - a) / a;
            mean_rel += rel; rms += rel * rel;
        }
        mean_rel /= 128 * 64; rms = sqrt(rms / (128 * 64));
        printf("PB chain of %2d MMAs (K=16 each): mean signed rel err %+.3e (= %+.2f x 2^-24 per MMA), rms %.3e\n", chain, mean_rel, mean_rel / chain * 16777216.0, rms);
    }
    printf("%s (%d failing cases)\n", fails ? "PROBE FAILED" : "PROBE OK", fails);
    return fails ? 1 : 0;
}
