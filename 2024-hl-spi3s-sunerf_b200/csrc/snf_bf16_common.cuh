// Constants and layouts shared by the bf16 tensor-core kernels (forward, dgrad chain, wgrad).
#pragma once
#include "snf_common.cuh"
#include "snf_tcgen05.cuh"

namespace snf {
namespace bf {
using namespace tc;

constexpr int TILE_M = 128;                  // points per CTA tile (= TMEM lanes)
constexpr int D = 512;                       // hidden width
constexpr int NH = 8;                        // hidden layers
constexpr int K0 = 96;                       // layer-0 K: 84 features + 4 bf16 residuals of x + 8 zero columns
constexpr int SLAB_BYTES = TILE_M * 128;     // one K-slab (64 bf16) of a 128-row image: 16 KB
constexpr int A_BYTES = 8 * SLAB_BYTES;      // 128 KB activation image
constexpr int WBLK_ROWS = 256;               // weight block: 256 output features x 64 k
constexpr int WBLK_BYTES = WBLK_ROWS * 128;  // 32 KB
constexpr int NSTAGE = 3;                    // weight ring depth (x 32 KB)
constexpr int N_EPI_WARPS = 8;
constexpr int N_EPI = N_EPI_WARPS * 32;      // 256 epilogue threads: (row, column half)
constexpr int NTHREADS = 64 + N_EPI;         // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int BIAS_BYTES = D * 4;
// shared memory: [A image 128 KB][weight ring 96 KB][bias 2 KB][barriers 128 B]; the base must be 1024-aligned
constexpr int SMEM_BYTES = A_BYTES + NSTAGE * WBLK_BYTES + BIAS_BYTES + 128;

constexpr int FWD_BLOCKS = 4 + 7 * 16;       // forward weight blocks: layer 0 (2 n-halves x 2 k-slabs) + 7 x 16
constexpr int WT_BLOCKS = 7 * 16;            // W^T blocks for the dgrad chain (layers 1..7)
// packed buffer: [FWD_BLOCKS x 32 KB][bias 8x512 f32][W_out 2x512 f32][b_out 2 f32 + pad] | [WT_BLOCKS x 32 KB]
constexpr int64_t PACK_W_BYTES = (int64_t)FWD_BLOCKS * WBLK_BYTES;
constexpr int64_t PACK_BIAS_OFF = PACK_W_BYTES;
constexpr int64_t PACK_WOUT_OFF = PACK_BIAS_OFF + NH * D * 4;
constexpr int64_t PACK_BOUT_OFF = PACK_WOUT_OFF + 2 * D * 4;
constexpr int WOUT_BYTES = 2 * D * 4;        // 4 KB, rides through the weight ring as a pseudo-block
constexpr int64_t PACK_WT_OFF = (PACK_BOUT_OFF + 16 + 1023) / 1024 * 1024;
constexpr int64_t PACK_TOTAL_BYTES = PACK_WT_OFF + (int64_t)WT_BLOCKS * WBLK_BYTES;

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// saved-image workspace (training): [enc][H = sin(pre)][P = pre][D = dL/dpre], all [tile][layer][128 KB] bf16 images
struct Bf16Ws {
  uint8_t *enc, *h, *pre, *d;
  int64_t bytes;
};
inline Bf16Ws bf16_layout(void *base, int64_t M, int train) {
  Bf16Ws w{};
  const int64_t tiles = (M + TILE_M - 1) / TILE_M;
  uint8_t *p = reinterpret_cast<uint8_t *>(base);
  int64_t off = 0;
  if (train) {
    w.enc = p + off; off += tiles * 2 * SLAB_BYTES;
    w.h = p + off; off += tiles * NH * (int64_t)A_BYTES;
    w.pre = p + off; off += tiles * NH * (int64_t)A_BYTES;
    w.d = p + off; off += tiles * NH * (int64_t)A_BYTES;
  }
  w.bytes = off > 0 ? off : 256;
  return w;
}

}  // namespace bf
}  // namespace snf
