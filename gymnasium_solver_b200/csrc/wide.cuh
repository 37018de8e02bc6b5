// wide.cuh — what the 256-wide tensor-core kernels (update_wide.cu, collect_wide.cu) share: tile / ring geometry, the cp.async.bulk helpers
// and the kernel that stages W2 in the operand layout.  See update_wide.cu for the design.
#pragma once

#include "mlp_tile.cuh"
#include "f16x3.cuh"
#include "update_shared.cuh"

namespace gs {

using namespace tc;

namespace wfu {

using namespace hfu;

constexpr int H = 256;
constexpr int kSlabs = 4;
constexpr uint32_t kTile = kSlabs * kSlab;          // one precision of a [128][256] tile: 64 KB
constexpr uint32_t kTileBytes = 4 * kTile;          // scratch per tile: h1 hi, h1 lo, dz2 hi, dz2 lo
constexpr uint32_t kW2Prec = H * H * 2;             // one precision of the staged W2: 128 KB
constexpr uint32_t kStage = 16384;
constexpr int kRing = 3;
constexpr int kStagesPerTile = 32;                  // 16 forward + 16 dgrad
constexpr int kFlushTiles = 64;

constexpr int kCW = 16;                             // compute warps
constexpr int kCompute = kCW * 32;
constexpr int kWideThreads = kCompute + 64;             // + MMA warp + loader warp

enum { BAR_Z1 = 0, BAR_Z2, BAR_OUT, BAR_DH2, BAR_WC, BAR_DH1, BAR_WB, RDY_X, RDY_H1, RDY_H2, RDY_G, RDY_DZ2, RDY_DZ1, BAR_FULL, BAR_EMPTY = BAR_FULL + kRing,
       kBars = BAR_EMPTY + kRing };

// shared memory
constexpr uint32_t oPhi = 0, oPlo = kTile, oX = 2 * kTile, oWS = oX + kSlab, oRing = oWS + H * 128, oBars = oRing + kRing * kStage;
constexpr uint32_t oTmem = oBars + 8 * kBars, oBH = oTmem + 16, oNcs = oBH + 16, oRed = oNcs + 32, oFr = oRed + 8 * PM_N, kSmemBytes = oFr + 64;
static_assert(kSmemBytes <= 232448, "shared memory budget");
// TMEM columns
constexpr uint32_t cAcc = 0, cH = 256, cW1 = 272, cWh = 304, cB2 = 336, cAhi = 368;   // cAhi: the hi half of the activation tile as TMEM A operand (128 columns)

// wgrad kernel: ring of 32-sample slices of the four stored tiles
constexpr int kGRing = 3;
constexpr uint32_t kGPiece = 32 * 128;              // 32 rows of one slab
constexpr uint32_t kGStage = 16 * kGPiece;          // {h1 hi, h1 lo, dz2 hi, dz2 lo} x 4 slabs = 64 KB
constexpr uint32_t kGSmem = kGRing * kGStage + 256;
enum { G_FULL = 0, G_EMPTY = kGRing, G_ACC = 2 * kGRing, G_FOLDED, kGBars };

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

}  // namespace wfu

// W2 (rows j, columns k) and W2^T (rows k, columns j), each as hi then lo, in the operand layout: staged[4][128 KB] = {W2 hi, W2 lo, W2^T hi,
// W2^T lo} (update_wide.cu)
int launch_stage_w2(const MlpDev& md, unsigned char* staged, cudaStream_t st);

}  // namespace gs
