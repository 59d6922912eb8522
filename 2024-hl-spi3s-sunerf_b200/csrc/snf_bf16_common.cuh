// Constants, layouts and role helpers shared by the bf16 tensor-core kernels (forward, dgrad chain, wgrad).
#pragma once
#include "snf_common.cuh"
#include "snf_tcgen05.cuh"

namespace snf {
namespace bf {
using namespace tc;

constexpr int TILE_M = 128;                  // points per CTA tile (= TMEM lanes); a CTA pair covers 256
constexpr int D = 512;                       // hidden width
constexpr int NH = 8;                        // hidden layers
constexpr int K0 = 96;                       // layer-0 K: 84 features + 4 bf16 residuals of x + 8 zero columns
constexpr int SLAB_BYTES = TILE_M * 128;     // one K-slab (64 bf16) of a 128-row image: 16 KB
constexpr int A_BYTES = 8 * SLAB_BYTES;      // 128 KB activation image

// ---- layer-chain kernels (forward, dgrad): CTA pairs, tcgen05 cta_group::2, M=256 N=128 K=16
constexpr int NCHUNK = 128;                  // output features per MMA / per weight block
constexpr int WBLK_BYTES = NCHUNK * 128;     // weight block: 128 output features x 64 k = 16 KB
constexpr int WHALF_BYTES = WBLK_BYTES / 2;  // each CTA of the pair streams half of every block: 8 KB
constexpr int NSTAGE = 8;                    // weight ring depth (x 8 KB per CTA)
constexpr int N_EPI_WARPS = 8;
constexpr int N_EPI = N_EPI_WARPS * 32;      // 256 epilogue threads: (row, column half)
constexpr int NTHREADS = 64 + N_EPI;         // warp 0 TMA producer, warp 1 MMA issuer / peer relay, warps 2..9 epilogue
constexpr int STG_WARP_BYTES = 32 * 128;     // per-warp staging: 32 rows of one 64-column slab = 4 KB (contiguous in an image)
constexpr int STG_BYTES = N_EPI_WARPS * STG_WARP_BYTES;
constexpr int BIAS_BYTES = D * 4;
// shared memory: [A image 128 KB][weight ring 64 KB][staging 32 KB][bias 2 KB][barriers]; base must be 1024-aligned
constexpr int OFF_RING = A_BYTES;
constexpr int OFF_STG = OFF_RING + NSTAGE * WHALF_BYTES;
constexpr int OFF_BIAS = OFF_STG + STG_BYTES;
constexpr int OFF_BAR = OFF_BIAS + BIAS_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;

constexpr int FWD_BLOCKS = 4 * 2 + 7 * 32;   // forward weight blocks: layer 0 (4 chunks x 2 k-slabs) + 7 x (4 x 8)
constexpr int WT_BLOCKS = 7 * 32;            // W^T blocks for the dgrad chain (layers 1..7)
// packed buffer: [FWD_BLOCKS x 16 KB][bias 8x512 f32][W_out 2x512 f32][b_out 2 f32 + pad] | [WT_BLOCKS x 16 KB]
constexpr int64_t PACK_W_BYTES = (int64_t)FWD_BLOCKS * WBLK_BYTES;
constexpr int64_t PACK_BIAS_OFF = PACK_W_BYTES;
constexpr int64_t PACK_WOUT_OFF = PACK_BIAS_OFF + NH * D * 4;
constexpr int64_t PACK_BOUT_OFF = PACK_WOUT_OFF + 2 * D * 4;
constexpr int WOUT_BYTES = 2 * D * 4;        // 4 KB, rides through the weight ring as a pseudo-block
constexpr int64_t PACK_WT_OFF = (PACK_BOUT_OFF + 16 + 1023) / 1024 * 1024;
constexpr int64_t PACK_TOTAL_BYTES = PACK_WT_OFF + (int64_t)WT_BLOCKS * WBLK_BYTES;

__host__ __device__ constexpr int fwd_layer_blocks(int l) { return l == 0 ? 8 : 32; }
__host__ __device__ constexpr int fwd_layer_first_block(int l) { return l == 0 ? 0 : 8 + (l - 1) * 32; }

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Barrier block (8 bytes each) inside [OFF_BAR, OFF_BAR + 256)
struct Bars {
  uint32_t base;
  __device__ uint32_t full(int s) const { return base + 8u * s; }                       // local TMA landed
  __device__ uint32_t empty(int s) const { return base + 8u * (NSTAGE + s); }           // stage consumed (multicast commit)
  __device__ uint32_t peer_full(int s) const { return base + 8u * (2 * NSTAGE + s); }   // leader only: peer's half landed
  __device__ uint32_t acc() const { return base + 8u * (3 * NSTAGE); }                  // layer accumulated (multicast commit)
  __device__ uint32_t aready() const { return base + 8u * (3 * NSTAGE + 1); }           // leader only: both A images ready
  __device__ uint32_t tmem_slot() const { return base + 8u * (3 * NSTAGE + 2); }
};

// saved-image workspace (training): [enc][H = sin(pre)][P = pre][D = dL/dpre], all [tile][layer][128 KB] bf16 images.
// Sized for an even number of tiles (CTA pairs always process two).
struct Bf16Ws {
  uint8_t *enc, *h, *pre, *d;
  int64_t bytes;
};
inline Bf16Ws bf16_layout(void *base, int64_t M, int train) {
  Bf16Ws w{};
  int64_t tiles = (M + TILE_M - 1) / TILE_M;
  tiles = (tiles + 1) / 2 * 2;
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
