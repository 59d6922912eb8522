// Split-precision tensor-core forward of the field network (sm_100a): the EXACT mode on tcgen05.
//
// north_star asks 1e-5 on rendered intensities in fp32 mode; through eight sine layers and an exp head neither bf16
// nor fp16 nor TF32 operands hold that (16-bit mode: 0.7-1.8e-5).  Round 1 met the gate with an FFMA SIMT SGEMM at
// 35 TFLOP/s.  Here every operand is the PAIR (hi, lo) = (fp16(v), fp16(v - hi)) - 22 significant bits - and a product
// is three tensor-core instructions accumulating into the same fp32 TMEM tile:  A_hi B_hi + A_hi B_lo + A_lo B_hi
// (the lo x lo term is below fp32 resolution).  Replaces NeRF.forward / NeRF_DT.forward (sunerf/model/model.py:44-57,
// 169-187) and PositionalEncoding.forward (:123-132) in "fp32 mode" for the default 8 x 512 network.
//
// The activation pair does not fit next to the weight ring in shared memory (2 x 128 KB), so unlike the fused 16-bit
// chain (snf_mlp_bf16.cu) this is ONE KERNEL PER LAYER over HBM-resident tile images:
//   work item  = (pair of 128-point tiles, N-half of 256 output features) per persistent CTA PAIR (cluster of 2, tcgen05
//                cta_group::2): CTA r owns tile 2 tp + r - its A operand and its 128 accumulator lanes - and supplies
//                half of every weight block, so a k-slab costs each SM 64 KB of L2 reads instead of the 96 KB of a
//                single-CTA tile (the first version: L2-bound at 45 % of the tensor rate)
//   warp 0     TMA producer: per k-slab one stage = A_hi, A_lo slabs (2 x 16 KB) + this CTA's halves of the W_hi, W_lo
//              blocks (2 x 16 KB) = 64 KB, three stages
//   warp 1     MMA issuer (leader CTA; the peer's warp 1 relays "my half of the stage has landed"): M = 256 across the
//              pair, N = 256, K = 16 per instruction, 3 instructions per k-step; accumulators double buffered in TMEM
//              (2 x 256 columns), so the epilogue of item i runs under the MMAs of item i + 1
//   warps 4-11 epilogue: TMEM -> + bias -> sin / cos -> (hi, lo) pairs -> the next layer's tile images; in training also
//              the saved fp16 activation image and the one-byte cosine code the 16-bit backward reads
// sin / cos: two-constant Cody-Waite reduction to [-pi, pi] in fp32, then MUFU (absolute error 4e-7 there; the MUFU's
// own range reduction is what loses accuracy at large arguments).  Output layer: per-thread partial dot products with
// W_out in fp32, combined in a fixed order by a small kernel (deterministic).
//
// The BACKWARD of this mode is the 16-bit one (snf_mlp_bf16_bwd.cu) on the high halves saved here, with the W^T
// operand of the dgrad chain split in two (hi, lo): per-parameter gradients 1e-4 of the oracle's (gate 1e-3).
#define SNF_EPI_GROUPS 2
#include "snf_bf16_common.cuh"

namespace snf {
namespace bf {
namespace x3 {

constexpr int NST = 3;
constexpr int STAGE_A = 2 * SLAB_BYTES;          // A_hi, A_lo of this CTA's tile
constexpr int STAGE_B = 2 * WHALF_BYTES;         // this CTA's halves (128 of 256 output features x 64 k) of W_hi, W_lo
constexpr int STAGE_BYTES = STAGE_A + STAGE_B;   // 64 KB
constexpr int OFF_BIAS = NST * STAGE_BYTES;      // [512] floats
constexpr int OFF_WOUT = OFF_BIAS + D * 4;       // [2][512] floats
constexpr int OFF_BAR = OFF_WOUT + 2 * D * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int N_EPI_W = 8;
constexpr int THREADS = 128 + N_EPI_W * 32;
static_assert(SMEM_BYTES <= 232448, "x3 forward exceeds the shared-memory window");

struct Params {
  const uint8_t *a_hi, *a_lo;          // input tile images of this layer: [tile][nslab][16 KB]
  int64_t a_hi_stride, a_lo_stride;    // bytes between tiles
  int nslab, k16_last;                 // k-slabs per tile (2: encoder image, 8: hidden) and k-steps of the last slab
  const uint8_t *w_hi, *w_lo;          // this layer's weight blocks: [(n-half, k-slab)][32 KB]
  const float *bias;                   // [512]
  uint8_t *o_hi, *o_lo;                // output tile images [tile][8][16 KB]; o_lo null: not needed (last layer)
  int64_t o_hi_stride, o_lo_stride;
  uint8_t *codes;                      // training: cosine codes [tile][64 KB] (C_BYTES layout), else null
  int64_t codes_stride;
  const float *w_out;                  // last layer: [2][512], else null
  float2 *part;                        // last layer: [tiles * 128][4] partial outputs
  int num_tiles;                       // even
};

// ---- positional encoding of [tiles * 128] rows (rows >= M encode x = 0) as the (hi, lo) layer-0 operand images.
// Columns: 0-3 x, 4-43 sin, 44-83 cos (f major, c minor), 84-87 the part of x that fp16(x) drops (it meets the
// raw-coordinate weights again, as in the 16-bit kernel), 88-127 zero.  One thread per (row, 8-column chunk).
__global__ void __launch_bounds__(256) x3_encode_kernel(const float4 *__restrict__ x, int64_t M, int64_t rows,
                                                        uint8_t *__restrict__ enc_hi, uint8_t *__restrict__ enc_lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 16) return;
  const int64_t m = idx >> 4;
  const int c8 = (int)(idx & 15);
  float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m < M) xv = x[m];
  const float xc[4] = {xv.x, xv.y, xv.z, xv.w};
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = c8 * 8 + i;
    float r = 0.f;
    if (col < 4) {
      r = __half2float(__float2half_rn(xc[col]));                    // exactly representable: its low half is zero
    } else if (col < 84) {
      const int e = (col - 4) % 40, f = e >> 2, c = e & 3;
      const float arg = fdiv(fmul(xc[c], (float)(1 << f)), 2.f);     // model.py:129, both scalings exact
      r = col < 44 ? sinf(arg) : cosf(arg);
    } else if (col < 88) {
      r = xc[col - 84] - __half2float(__float2half_rn(xc[col - 84]));
    }
    v[i] = r;
  }
  uint4 hi, lo;
  float res[8];
  uint32_t *hp = reinterpret_cast<uint32_t *>(&hi), *lp = reinterpret_cast<uint32_t *>(&lo);
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const __half2 h2 = __floats2half2_rn(v[i], v[i + 1]);
    const float2 back = __half22float2(h2);
    res[i] = v[i] - back.x; res[i + 1] = v[i + 1] - back.y;
    hp[i / 2] = *reinterpret_cast<const uint32_t *>(&h2);
    lp[i / 2] = pack_f16x2(res[i], res[i + 1]);
  }
  const int64_t tile = m / TILE_M;
  const int row = (int)(m % TILE_M);
  const int64_t off = tile * 2 * SLAB_BYTES + (c8 >> 3) * SLAB_BYTES + sw128_chunk_off(row, c8 & 7);
  *reinterpret_cast<uint4 *>(enc_hi + off) = hi;
  *reinterpret_cast<uint4 *>(enc_lo + off) = lo;
}

// sin and cos of v: reduce to r in [-pi, pi] with a two-constant Cody-Waite step (exact for |v| < 2^11 pi), MUFU on r
__device__ __forceinline__ void sincos_reduced(float v, float &s, float &c) {
  const float k = rintf(v * 0.15915494309189535f);
  float r = fmaf(k, -6.28318548202514648f, v);        // 2 pi rounded to float
  r = fmaf(k, 1.74845553146951715e-7f, r);            // float(2 pi) - 2 pi
  s = __sinf(r);
  c = __cosf(r);
}
// the backward's one-byte cosine code (snf_bf16_common.cuh) from the cosine itself
__device__ __forceinline__ uint32_t cosq_enc_full(float c) {
  const uint32_t q = __float_as_uint(fmaf(sqrtf(1.f - fabsf(c)), COSQ_T, COSQ_MAGIC));
  return c < 0.f ? (q | 0x80u) : q;
}

// chunks c8, c8 + 1 (c8 even; v[0..3], v[4..7]) of row `row` of a SWIZZLE_128B slab as ONE 32-byte store
__device__ __forceinline__ void st_row_pair(uint8_t *slab, int row, int c8, const uint32_t *v) {
  const bool odd = row & 1;                       // (c8 ^ r) and (c8 + 1) ^ r share the aligned pair; odd rows swap them
  uint8_t *dst = slab + (row >> 3) * 1024 + (row & 7) * 128 + (((c8 ^ (row & 7)) & ~1) << 4);
  const uint32_t a0 = odd ? v[4] : v[0], a1 = odd ? v[5] : v[1], a2 = odd ? v[6] : v[2], a3 = odd ? v[7] : v[3];
  const uint32_t b0 = odd ? v[0] : v[4], b1 = odd ? v[1] : v[5], b2 = odd ? v[2] : v[6], b3 = odd ? v[3] : v[7];
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0),
               "r"(b1), "r"(b2), "r"(b3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) mlp_x3_layer_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  float *bias_s = reinterpret_cast<float *>(smem_raw + OFF_BIAS);
  float *wout_s = reinterpret_cast<float *>(smem_raw + OFF_WOUT);
  const uint32_t sBar = base + OFF_BAR;
  auto bar_full = [&](int s) { return sBar + 8u * s; };                    // leader: own TMA + the peer's relay
  auto bar_empty = [&](int s) { return sBar + 8u * (NST + s); };           // multicast tcgen05.commit
  auto bar_accf = [&](int j) { return sBar + 8u * (2 * NST + j); };        // accumulator j complete (multicast commit)
  auto bar_acce = [&](int j) { return sBar + 8u * (2 * NST + 2 + j); };    // leader: accumulator j drained by the 16 epilogue warps of the pair
  const uint32_t tmem_slot = sBar + 8u * (2 * NST + 4);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full(s), rank == 0 ? 2 : 1); mbar_init(bar_empty(s), 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(bar_accf(j), 1); mbar_init(bar_acce(j), 2 * N_EPI_W); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < D; i += THREADS) bias_s[i] = __ldg(p.bias + i);
  if (p.w_out != nullptr)
    for (int i = threadIdx.x; i < 2 * D; i += THREADS) wout_s[i] = __ldg(p.w_out + i);
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem_raw + OFF_BAR + 8 * (2 * NST + 4));
  const int n_items = p.num_tiles;                     // (tile pairs) x (2 N-halves); item -> (tile pair item >> 1, N-half item & 1)

  if (warp == 0) {
    if (lane == 0) {
      // =========================== TMA producer: this CTA's tile and its half of every weight block
      const uint64_t keep = l2_policy_evict_last();
      int s = 0; uint32_t ph = 0;
      for (int item = pair; item < n_items; item += npairs) {
        const int64_t tile = (int64_t)(item >> 1) * 2 + rank;
        const int nh = item & 1;
        const uint8_t *ah = p.a_hi + tile * p.a_hi_stride, *al = p.a_lo + tile * p.a_lo_stride;
        for (int ks = 0; ks < p.nslab; ++ks) {
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_arrive_expect_tx(bar_full(s), STAGE_BYTES);
          const uint32_t st = base + s * STAGE_BYTES;
          const int64_t wb = (int64_t)(nh * p.nslab + ks) * WBLK_BYTES + rank * WHALF_BYTES;
          bulk_g2s(st, ah + (int64_t)ks * SLAB_BYTES, SLAB_BYTES, bar_full(s));
          bulk_g2s(st + SLAB_BYTES, al + (int64_t)ks * SLAB_BYTES, SLAB_BYTES, bar_full(s));
          bulk_g2s_hint(st + STAGE_A, p.w_hi + wb, WHALF_BYTES, bar_full(s), keep);
          bulk_g2s_hint(st + STAGE_A + WHALF_BYTES, p.w_lo + wb, WHALF_BYTES, bar_full(s), keep);
          if (++s == NST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    if (rank == 0) {
      // =========================== MMA issuer (leader CTA): uniform control flow, one elected lane issues
      const uint32_t idesc = idesc_f16kind(2 * TILE_M, NCHUNK, FMT_F16, FMT_F16);
      uint32_t eph = 0;
      int it = 0;
      for (int item = pair; item < n_items; item += npairs, ++it) {
        const int j = it & 1;
        if (it >= 2) { mbar_wait(bar_acce(j), (eph >> j) & 1u); eph ^= 1u << j; }   // both epilogues have drained this buffer
        tcgen05_fence_after();
        const uint32_t dcol = tmem + j * NCHUNK;
        for (int ks = 0; ks < p.nslab; ++ks) {
          mbar_wait(bar_full(s), ph);                    // both CTAs' parts of the stage have landed
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t st = base + s * STAGE_BYTES;
            const uint64_t ahi = smem_desc(st, 16, 1024), alo = smem_desc(st + SLAB_BYTES, 16, 1024);
            const uint64_t bhi = smem_desc(st + STAGE_A, 16, 1024), blo = smem_desc(st + STAGE_A + WHALF_BYTES, 16, 1024);
            const int nk = ks == p.nslab - 1 ? p.k16_last : 4;
            for (int k4 = 0; k4 < nk; ++k4) {
              mma_ss_2cta(dcol, ahi + 2 * k4, bhi + 2 * k4, idesc, (ks | k4) != 0);
              mma_ss_2cta(dcol, ahi + 2 * k4, blo + 2 * k4, idesc, 1);
              mma_ss_2cta(dcol, alo + 2 * k4, bhi + 2 * k4, idesc, 1);
            }
            mma_commit_2cta(bar_empty(s), 3);            // frees the stage in both CTAs
            if (ks == p.nslab - 1) mma_commit_2cta(bar_accf(j), 3);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1; }
        }
      }
    } else if (lane == 0) {
      // =========================== peer relay: tell the leader when this CTA's part of a stage has landed
      for (int item = pair; item < n_items; item += npairs)
        for (int ks = 0; ks < p.nslab; ++ks) {
          mbar_wait(bar_full(s), ph);
          mbar_arrive_remote_relaxed(mapa_shared(bar_full(s), 0));   // the data is TMA-written and tensor-core-read
          if (++s == NST) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp >= 4) {
    // =========================== epilogue: thread = (row of this CTA's tile, column half g of the 256-column accumulator)
    const int e = warp - 4, q = warp & 3, g = e >> 2;
    const int row = q * 32 + lane;
    uint32_t fph = 0;
    int it = 0;
    const uint32_t acce_addr[2] = {rank == 0 ? bar_acce(0) : mapa_shared(bar_acce(0), 0),
                                   rank == 0 ? bar_acce(1) : mapa_shared(bar_acce(1), 0)};
    for (int item = pair; item < n_items; item += npairs, ++it) {
      const int j = it & 1;
      const int64_t tile = (int64_t)(item >> 1) * 2 + rank;
      const int nh = item & 1;
      mbar_wait(bar_accf(j), (fph >> j) & 1u); fph ^= 1u << j;
      tcgen05_fence_after();
      const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16) + j * NCHUNK + g * 128;
      uint8_t *ohi = p.o_hi != nullptr ? p.o_hi + tile * p.o_hi_stride : nullptr;
      uint8_t *olo = p.o_lo != nullptr ? p.o_lo + tile * p.o_lo_stride : nullptr;
      uint8_t *cod = p.codes != nullptr ? p.codes + tile * p.codes_stride : nullptr;
      float o0 = 0.f, o1 = 0.f;
      uint32_t accA[32], accB[32];
      tmem_ld32(tm_row, accA);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? accB : accA;
        uint32_t(&nxt)[32] = (c & 1) ? accA : accB;
        tmem_ld_wait(cur);
        if (c + 1 < 4) tmem_ld32(tm_row + (c + 1) * 32, nxt);
        const int ncol = nh * NCHUNK + g * 128 + c * 32;   // column of the 512-wide layer
        const int slab = ncol >> 6, c8_0 = (ncol & 63) >> 3;
        uint32_t hi[16], lo[16], cq[8];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = *reinterpret_cast<const float4 *>(bias_s + ncol + i);
          const float bb[4] = {b.x, b.y, b.z, b.w};
          float sv[4], cv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) sincos_reduced(__uint_as_float(cur[i + k]) + bb[k], sv[k], cv[k]);
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            const __half2 h2 = __floats2half2_rn(sv[k], sv[k + 1]);
            const float2 back = __half22float2(h2);
            hi[(i + k) / 2] = *reinterpret_cast<const uint32_t *>(&h2);
            lo[(i + k) / 2] = pack_f16x2(sv[k] - back.x, sv[k + 1] - back.y);
          }
          if (cod != nullptr) cq[i / 4] = cosq_pack4(cosq_enc_full(cv[0]), cosq_enc_full(cv[1]), cosq_enc_full(cv[2]), cosq_enc_full(cv[3]));
          if (p.w_out != nullptr) {
            const float4 wa = *reinterpret_cast<const float4 *>(wout_s + ncol + i);
            const float4 wb = *reinterpret_cast<const float4 *>(wout_s + D + ncol + i);
            o0 += sv[0] * wa.x + sv[1] * wa.y + sv[2] * wa.z + sv[3] * wa.w;
            o1 += sv[0] * wb.x + sv[1] * wb.y + sv[2] * wb.z + sv[3] * wb.w;
          }
        }
        // Each lane owns a ROW of the tile image (rows are 128 B apart), so a warp store touches 32 lines whatever its
        // width: the kernel was bound by these L1 wavefronts (8192 per item in rendering against 12288 MMA cycles - the 67 %
        // tensor-pipe activity ncu showed).  sm_100 has 256-bit stores: the 128-byte swizzle XORs the chunk index with
        // row mod 8, which keeps an aligned PAIR of 16-byte chunks together (swapped in odd rows) - half the instructions.
        if (ohi != nullptr) {
#pragma unroll
          for (int k = 0; k < 4; k += 2) st_row_pair(ohi + slab * SLAB_BYTES, row, c8_0 + k, &hi[4 * k]);
        }
        if (olo != nullptr) {
#pragma unroll
          for (int k = 0; k < 4; k += 2) st_row_pair(olo + slab * SLAB_BYTES, row, c8_0 + k, &lo[4 * k]);
        }
        if (cod != nullptr) {
          const int ch16 = (ncol & 63) >> 4;
#pragma unroll
          for (int k = 0; k < 2; ++k)
            __stcs(reinterpret_cast<uint4 *>(cod + (((slab * 4 + ch16 + k) * TILE_M + row) << 4)),
                   make_uint4(cq[4 * k], cq[4 * k + 1], cq[4 * k + 2], cq[4 * k + 3]));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(acce_addr[j]);
        else mbar_arrive_remote_relaxed(acce_addr[j]);
      }
      if (p.part != nullptr) p.part[(tile * TILE_M + row) * 4 + nh * 2 + g] = make_float2(o0, o1);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

// out[m] = ((p0 + p1) + (p2 + p3)) + b_out + offset: the four partial dot products of a point in a fixed order
__global__ void __launch_bounds__(256) x3_out_kernel(const float2 *__restrict__ part, int64_t M, const float *__restrict__ b_out,
                                                     float off0, float off1, float2 *__restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float4 a = *reinterpret_cast<const float4 *>(part + m * 4), b = *reinterpret_cast<const float4 *>(part + m * 4 + 2);
  out[m] = make_float2(((a.x + a.z) + (b.x + b.z)) + __ldg(b_out) + off0, ((a.y + a.w) + (b.y + b.w)) + __ldg(b_out + 1) + off1);
}

}  // namespace x3
}  // namespace bf
}  // namespace snf

using namespace snf;

int snf_x3_set_attributes() {
  return (int)cudaFuncSetAttribute(bf::x3::mlp_x3_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::x3::SMEM_BYTES);
}

extern "C" int snf_mlp_fwd_x3(const float *x, int64_t M, const void *packed, float off0, float off1, float *out, void *ws,
                              int train, void *stream) {
  if (M == 0) return 0;
  SNF_CHECK_PTR(x); SNF_CHECK_PTR(packed); SNF_CHECK_PTR(out); SNF_CHECK_PTR(ws);
  SNF_CHECK_ALIGN(x, 16); SNF_CHECK_ALIGN(out, 8); SNF_CHECK_ALIGN(packed, 1024); SNF_CHECK_ALIGN(ws, 1024);
  if (M < 0) return SNF_E_ARG;
  int nsm = 0;
  if (int e = snf_device_setup(&nsm)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const bf::Bf16Ws w = bf::bf16_layout(ws, M, train, 1);
  int num_tiles = (int)((M + bf::TILE_M - 1) / bf::TILE_M);
  num_tiles = (num_tiles + 1) / 2 * 2;                 // the 16-bit backward works on pairs of tiles: fill the padding tile too
  const int64_t rows = (int64_t)num_tiles * bf::TILE_M;
  const uint8_t *pk = reinterpret_cast<const uint8_t *>(packed);
  bf::x3::x3_encode_kernel<<<(unsigned)ceil_div64(rows * 16, 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(x), M, rows,
                                                                                w.enc, w.enc_lo);
  int grid = num_tiles < nsm ? num_tiles : nsm;     // CTA pairs: (num_tiles / 2) tile pairs x 2 N-halves = num_tiles items per pair slot
  grid &= ~1;
  for (int l = 0; l < bf::NH; ++l) {
    bf::x3::Params p{};
    const int64_t img = bf::A_BYTES;
    if (l == 0) {
      p.a_hi = w.enc; p.a_lo = w.enc_lo; p.a_hi_stride = p.a_lo_stride = 2 * bf::SLAB_BYTES;
      p.nslab = 2; p.k16_last = (bf::K0 - 64) / 16;
      p.w_hi = pk; p.w_lo = pk + bf::PACK_LO_OFF;
    } else {
      if (train) { p.a_hi = w.h + (int64_t)(l - 1) * img; p.a_hi_stride = bf::NH * img; }
      else { p.a_hi = w.h_pp[(l - 1) & 1]; p.a_hi_stride = img; }
      p.a_lo = w.h_lo[(l - 1) & 1]; p.a_lo_stride = img;
      p.nslab = 8; p.k16_last = 4;
      const int64_t boff = (int64_t)(4 + (l - 1) * 16) * bf::WBLK_BYTES;
      p.w_hi = pk + boff; p.w_lo = pk + bf::PACK_LO_OFF + boff;
    }
    p.bias = reinterpret_cast<const float *>(pk + bf::PACK_BIAS_OFF) + l * bf::D;
    const bool last = l == bf::NH - 1;
    if (train) { p.o_hi = w.h + (int64_t)l * img; p.o_hi_stride = bf::NH * img; }
    else if (!last) { p.o_hi = w.h_pp[l & 1]; p.o_hi_stride = img; }
    if (!last) { p.o_lo = w.h_lo[l & 1]; p.o_lo_stride = img; }
    if (train) { p.codes = w.pre + (int64_t)l * bf::C_BYTES; p.codes_stride = (int64_t)bf::NH * bf::C_BYTES; }
    if (last) { p.w_out = reinterpret_cast<const float *>(pk + bf::PACK_WOUT_OFF); p.part = w.part; }
    p.num_tiles = num_tiles;
    bf::x3::mlp_x3_layer_kernel<<<grid, bf::x3::THREADS, bf::x3::SMEM_BYTES, st>>>(p);
  }
  bf::x3::x3_out_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, st>>>(w.part, M, reinterpret_cast<const float *>(pk + bf::PACK_BOUT_OFF),
                                                                     off0, off1, reinterpret_cast<float2 *>(out));
  count_launch(2 + bf::NH);
  return launch_status();
}
