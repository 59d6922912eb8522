// Inference forward of the bf16 field network with the ACTIVATION OPERAND IN TENSOR MEMORY (tcgen05.mma "TS" form).
// An alternative to the shared-memory-operand kernel of snf_mlp_bf16.cu, selected with snf_debug_fwd_variant(1); NOT the
// default: it is bit-compatible (same accumulation order; outputs differ from the SS kernel by the summation order of the
// 512 -> 2 output layer only, ~2e-8) but measured 5-12 % slower.
//
// Why it was built: the SS kernel is bound by shared-memory bandwidth - per half layer an SM moves 32 x 12 KB of operand
// fetch, 128 KB of weight stages and 64 KB of epilogue stores against 4.1 k cycles of tensor time.  Here the activation
// operand lives in TMEM (128 rows x 512 bf16 = 256 columns, written by the epilogue with tcgen05.st), an instruction
// fetches only its weights from shared memory and the epilogue does not touch shared memory at all.
// Why it is not faster: with 256 of the 512 TMEM columns taken by the operand the accumulator shrinks to two 128-column
// buffers, i.e. N = 128 per instruction, and an N = 128 instruction re-reads its 4 KB of A every 64 cycles - 64 B/clk, the
// whole TMEM read rate - while the epilogue reads the finished accumulators through the same port.  (Neither the sine
// (removing it changes nothing) nor the issue loop (3 SASS instructions per MMA, one synchronisation per quarter) is
// the limit.)
//
//   TMEM   columns [0,128) and [128,256): two accumulator buffers of one output QUARTER each (128 features, fp32),
//          columns [256,512): the activation operand of the current layer (column = k / 2).
//   MMA    cta_group::2, M = 256 (pair) x N = 128 x K = 16; a layer = 4 quarters x 32 instructions.  The epilogue of
//          quarter q runs under the MMAs of quarter q+1 (other buffer).
//   hold   the operand may only be overwritten where no remaining MMA of the layer reads it, so the finished quarters 0-2
//          wait in registers (96 per thread); quarter 3 is accumulated in ascending k and commits a barrier per k-block,
//          after which the held quarter of the same index is stored (under the remaining MMAs); the next layer's first
//          24 instructions (k < 384) queue up behind this layer's last one, the last 8 wait for quarter 3's epilogue.
//   layer 0 reads the positional encoding from a 32 KB shared-memory image (SS form, K = 96), as the SS kernel does.
//   weights: same TMA ring as the SS kernel, but stages of [64 features per CTA x 128 k] in (layer, quarter, k-block)
//          order (snf_mlp_pack_bf16 writes that order too when the variant is selected); 8 stages of 16 KB, a quarter
//          consumes 4; the issuer synchronises once per quarter (a wait / fence / commit per 8 instructions costs it
//          ~200 cycles and makes it the bound at 64 tensor cycles per instruction: tools/micro/mma_rate.cu).
// Same numerics as the SS kernel: fp32 accumulation, bf16 activations, sin.approx on fp32 pre-activations, output layer as
// an fp32 register dot product.
#define SNF_EPI_GROUPS 2
#include "snf_bf16_common.cuh"

namespace snf {
namespace bf {
namespace ts {

constexpr int NSTAGE = 8;
constexpr int ENC_BYTES = 2 * SLAB_BYTES;                    // layer-0 operand: 2 k-slabs of 128 rows
constexpr int OFF_RING = ENC_BYTES;
constexpr int OFF_BIAS = OFF_RING + NSTAGE * WHALF_BYTES;
constexpr int OFF_WOUT = OFF_BIAS + 2 * BIAS_BYTES;
constexpr int OFF_OSUM = OFF_WOUT + WOUT_BYTES;
constexpr int OFF_BAR = OFF_OSUM + TILE_M * 8;
constexpr int SMEM_BYTES = OFF_BAR + 512;
static_assert(SMEM_BYTES <= 232448, "shared-memory window");
constexpr int N_EPI_W = 8;                                   // epilogue warps: thread = (row, 64-column half of a quarter)
constexpr int THREADS = 128 + N_EPI_W * 32;

struct Bars {
  uint32_t base;
  __device__ uint32_t full(int s) const { return base + 8u * s; }
  __device__ uint32_t empty(int s) const { return base + 8u * (NSTAGE + s); }
  __device__ uint32_t acc(int b) const { return base + 8u * (2 * NSTAGE + b); }        // quarter accumulated in buffer b
  __device__ uint32_t dfree(int b) const { return base + 8u * (2 * NSTAGE + 2 + b); }  // leader: buffer b read out (16 warps)
  __device__ uint32_t ready(int k) const { return base + 8u * (2 * NSTAGE + 4 + k); }  // leader: 0 enc image, 1 operand k<384, 2 k>=384
  __device__ uint32_t tmem_slot() const { return base + 8u * (2 * NSTAGE + 7); }
  // kfree[kb]: the MMAs of quarter 3 over k-block kb are complete (multicast commit): no instruction of this layer reads
  // operand columns k < 128 (kb + 1) any more
  __device__ uint32_t kfree(int kb) const { return base + 8u * (2 * NSTAGE + 8 + kb); }
};
constexpr int TMEM_SLOT_OFF = OFF_BAR + 8 * (2 * NSTAGE + 7);

}  // namespace ts

// weight stages for the TS kernel: [block = (layer, quarter, k-block)][CTA rank][2 slabs][64 rows][128 B swizzled];
// CTA r supplies features 128 q + 64 r + row.  Layer 0 has one k-block per quarter (K0 = 96 of 128 columns used).
__global__ void __launch_bounds__(256) pack_weights_ts_kernel(const float *w0, const float *w1, const float *w2,
                                                              const float *w3, const float *w4, const float *w5,
                                                              const float *w6, const float *w7, uint4 *__restrict__ dst) {
  const float *W[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)FWD_BLOCKS * (WBLK_BYTES / 16)) return;
  const int blk = (int)(idx / (WBLK_BYTES / 16));
  const int within = (int)(idx % (WBLK_BYTES / 16));   // 2048 chunks: [rank][slab][row][pos]
  int l, q, kb;
  if (blk < 4) { l = 0; q = blk; kb = 0; }
  else { const int b2 = blk - 4; l = 1 + b2 / 16; q = (b2 % 16) >> 2; kb = b2 & 3; }
  const int rnk = within >> 10, slab = (within >> 9) & 1, r = (within >> 3) & 63, pos = within & 7;
  const int c8 = pos ^ (r & 7);
  const int n = q * 128 + rnk * 64 + r;
  const int kbase = kb * 128 + slab * 64 + c8 * 8;
  const int kin = l == 0 ? 84 : D;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = kbase + i;
    float x = 0.f;
    if (l == 0) {
      if (k < 84) x = W[0][n * kin + k];
      else if (k < 88) x = W[0][n * kin + (k - 84)];   // residual columns reuse the raw-coordinate weights
    } else {
      x = W[l][n * kin + k];
    }
    v[i] = x;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  dst[(int64_t)blk * (WBLK_BYTES / 16) + within] = o;
}

struct FwdTsParams {
  const float4 *x;
  int64_t M;
  int num_tiles;
  const uint8_t *packed;   // the packed buffer of snf_mlp_pack_bf16 (TS stages at PACK_TS_OFF)
  float2 *out;
  float off0, off1;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ts::THREADS, 1) mlp_fwd_ts_bf16_kernel(const FwdTsParams p) {
  using namespace ts;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  uint8_t *gA = smem_raw;                                                   // encoder image (layer-0 operand)
  const uint32_t sA = base, sW = base + OFF_RING;
  float *bias_s = reinterpret_cast<float *>(smem_raw + OFF_BIAS);           // [2][512]
  float *wout_s = reinterpret_cast<float *>(smem_raw + OFF_WOUT);           // [2][512]
  float2 *osum_s = reinterpret_cast<float2 *>(smem_raw + OFF_OSUM);         // [128]
  const Bars bar{base + OFF_BAR};
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *bias_all = reinterpret_cast<const float *>(p.packed + PACK_BIAS_OFF);
  const float *b_out = reinterpret_cast<const float *>(p.packed + PACK_BOUT_OFF);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar.full(s), rank == 0 ? 2 : 1); mbar_init(bar.empty(s), 1); }
    mbar_init(bar.acc(0), 1); mbar_init(bar.acc(1), 1);
    mbar_init(bar.dfree(0), 2 * N_EPI_W); mbar_init(bar.dfree(1), 2 * N_EPI_W);
    for (int k = 0; k < 3; ++k) mbar_init(bar.ready(k), 2 * N_EPI_W);
    for (int k = 0; k < 3; ++k) mbar_init(bar.kfree(k), 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * D; i += THREADS) wout_s[i] = __ldg(reinterpret_cast<const float *>(p.packed + PACK_WOUT_OFF) + i);
  for (int i = threadIdx.x; i < D; i += THREADS) bias_s[i] = __ldg(bias_all + i);
  if (warp == 1) tmem_alloc_2cta(bar.tmem_slot(), 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem_raw + TMEM_SLOT_OFF);

  if (warp < EPI_WARP0) {
    reg_dealloc<REGS_CTRL>();
    if (warp == 0) {
      // =========================== TMA producer: this CTA's 16 KB of every weight stage
      if (lane == 0) {
        int s = 0; uint32_t ph = 0;
        const uint64_t keep = l2_policy_evict_last();
        const uint8_t *src = p.packed + PACK_TS_OFF + rank * WHALF_BYTES;
        for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
          for (int blk = 0; blk < FWD_BLOCKS; ++blk) {
            mbar_wait(bar.empty(s), ph ^ 1);
            mbar_arrive_expect_tx(bar.full(s), WHALF_BYTES);
            bulk_g2s_hint(sW + s * WHALF_BYTES, src + (int64_t)blk * WBLK_BYTES, WHALF_BYTES, bar.full(s), keep);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      int s = 0; uint32_t ph = 0;
      if (rank == 0) {
        // =========================== MMA issuer (leader CTA), whole warp in uniform control flow, one elected lane issues
        const uint32_t idesc = idesc_bf16(256, 128);
        const uint64_t adesc0 = smem_desc(sA, 16, 1024), bdesc0 = smem_desc(sW, 16, 1024);
        uint32_t rph = 0, dph = 0;
        auto wait_ready = [&](int k) { mbar_wait(bar.ready(k), (rph >> k) & 1u); rph ^= 1u << k; };
        auto wait_dfree = [&](int b) { mbar_wait(bar.dfree(b), (dph >> b) & 1u); dph ^= 1u << b; };
        // One synchronisation (stage waits, tcgen05 fence, commits) costs the issuing thread ~200 cycles and an instruction
        // at least 47 (tools/micro/mma_rate.cu): at 64 tensor cycles per N = 128 instruction a wait / fence / commit per
        // 8 instructions makes the issuer the bound (89 cycles per instruction).  So the stages of a quarter are
        // waited for together and their "empty" commits follow the last instruction of the group.
        for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
          for (int l = 0; l < NH; ++l) {
            for (int q = 0; q < 4; ++q) {
              const int b = q & 1;
              wait_dfree(b);                                // the epilogue has read this buffer's previous quarter out
              const uint32_t d_tmem = tmem + 128 * b;
              if (l == 0) {
                if (q == 0) wait_ready(0);                  // encoder image of this tile
                mbar_wait(bar.full(s), ph);
                tcgen05_fence_after();
                if (elect_one()) {
                  const uint64_t bd = bdesc0 + (uint64_t)((s * WHALF_BYTES) >> 4);
#pragma unroll
                  for (int j = 0; j < K0 / 16; ++j)
                    mma_ss_2cta(d_tmem, adesc0 + (uint64_t)(((j >> 2) * SLAB_BYTES) >> 4) + 2 * (j & 3),
                                bd + (uint64_t)(((j >> 2) * (WHALF_BYTES / 2)) >> 4) + 2 * (j & 3), idesc, j != 0);
                  mma_commit_2cta(bar.empty(s), 3);
                  mma_commit_2cta(bar.acc(b), 3);
                }
                __syncwarp();
                if (++s == NSTAGE) { s = 0; ph ^= 1; }
                continue;
              }
              // layers >= 1.  A quarter consumes four consecutive ring slots starting at s = 0 or 4 (layer 0 used four
              // single-stage quarters), so every descriptor of its 32 instructions is one base + a compile-time offset:
              // the issue loop has to stay under 64 cycles per instruction.
              const uint64_t bdq = bdesc0 + (uint64_t)((s * WHALF_BYTES) >> 4);
              const uint32_t a0 = tmem + 256;
              auto issue = [&](int kb_lo, int kb_hi) {     // k-blocks [kb_lo, kb_hi), all bounds compile-time after inlining
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                  if (kb < kb_lo || kb >= kb_hi) continue;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    mma_ts_2cta(d_tmem, a0 + (kb * 8 + j) * 8,
                                bdq + (uint64_t)((kb * WHALF_BYTES + (j >> 2) * (WHALF_BYTES / 2)) >> 4) + 2 * (j & 3), idesc,
                                (kb | j) != 0);
                  if (q == 3 && kb < 3 && l < NH - 1) mma_commit_2cta(bar.kfree(kb), 3);   // the epilogue stores h_l early
                }
              };
              if (q == 0) {
                wait_ready(1);                              // operand columns k < 384 of h_{l-1}
                for (int i = 0; i < 3; ++i) mbar_wait(bar.full(s + i), ph);
                tcgen05_fence_after();
                if (elect_one()) issue(0, 3);
                __syncwarp();
                wait_ready(2);                              // k >= 384
                mbar_wait(bar.full(s + 3), ph);
                tcgen05_fence_after();
                if (elect_one()) issue(3, 4);
              } else {
                for (int i = 0; i < 4; ++i) mbar_wait(bar.full(s + i), ph);
                tcgen05_fence_after();
                if (elect_one()) issue(0, 4);
              }
              if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mma_commit_2cta(bar.empty(s + i), 3);
                mma_commit_2cta(bar.acc(b), 3);
              }
              __syncwarp();
              s += 4;
              if (s >= NSTAGE) { s -= NSTAGE; ph ^= 1; }
            }
          }
        }
      } else if (lane == 0) {
        for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
          for (int blk = 0; blk < FWD_BLOCKS; ++blk) {
            mbar_wait(bar.full(s), ph);
            mbar_arrive_remote_relaxed(mapa_shared(bar.full(s), 0));
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    reg_alloc<REGS_EPI>();
    // =========================== epilogue warps: thread = (row, g); in quarter q it owns the accumulator columns
    // 64 g .. 64 g + 63 of the quarter's buffer = features 128 q + 64 g + i = operand columns 256 + 64 q + 32 g + i / 2
    const int e = warp - EPI_WARP0;
    const int lq = warp & 3;                          // TMEM lane quarter this warp may access
    const int g = e >> 2;
    const int row = lq * 32 + lane;
    const int et = threadIdx.x - EPI_WARP0 * 32;
    const uint32_t tm_lane = tmem + ((uint32_t)(lq * 32) << 16);
    auto remote = [&](uint32_t a) { return rank == 0 ? a : mapa_shared(a, 0); };
    const uint32_t a_dfree[2] = {remote(bar.dfree(0)), remote(bar.dfree(1))};
    const uint32_t a_ready[3] = {remote(bar.ready(0)), remote(bar.ready(1)), remote(bar.ready(2))};
    auto arrive = [&](uint32_t addr) {                // one arrival per warp
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(addr);
        else mbar_arrive_remote_relaxed(addr);
      }
    };
    uint32_t aph = 0;                                 // phases of acc[0], acc[1]
    uint32_t kph = 0;                                 // phase of kfree[0..2] (one completion per layer >= 1)
    tcgen05_fence_before();
    arrive(a_dfree[0]); arrive(a_dfree[1]);           // both accumulator buffers start free
    for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
      const int tile = tp * 2 + (int)rank;
      const int64_t m = (int64_t)tile * TILE_M + row;
      // ---- layer-0 operand: positional encoding into the shared-memory image (slabs 0, 1), as in the SS kernel.
      //      (The image is only read by this tile's layer-0 MMAs, which completed long before the previous tile ended.)
      {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M) xv = p.x[m];
        const float xc[4] = {xv.x, xv.y, xv.z, xv.w};
        auto put8 = [&](int feat0, float a, float b, float c, float d) {
          const int c8 = feat0 >> 3;
          *reinterpret_cast<uint2 *>(gA + (c8 >> 3) * SLAB_BYTES + sw128_chunk_off(row, c8 & 7) + (feat0 & 7) * 2) =
              make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
        };
        if (g == 0) put8(0, xc[0], xc[1], xc[2], xc[3]);
        if (g == 1) {
          float rs[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) rs[c] = xc[c] - __bfloat162float(__float2bfloat16_rn(xc[c]));
          put8(84, rs[0], rs[1], rs[2], rs[3]);
          put8(88, 0.f, 0.f, 0.f, 0.f);
          put8(92, 0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int fi = 0; fi < 5; ++fi) {
          const int f = fi * 2 + g;
          float sv[4], cv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) sincosf(xc[c] * (float)(1 << f) * 0.5f, &sv[c], &cv[c]);
          put8(4 + f * 4, sv[0], sv[1], sv[2], sv[3]);
          put8(44 + f * 4, cv[0], cv[1], cv[2], cv[3]);
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        arrive(a_ready[0]);
      }
#pragma unroll 1
      for (int l = 0; l < NH; ++l) {
        const bool last = (l == NH - 1);
        const float *bl = bias_s + (l & 1) * D;
        uint32_t held[96];                            // quarters 0-2 of h_l, bf16 pairs
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int b = q & 1;
          mbar_wait(bar.acc(b), (aph >> b) & 1u);
          aph ^= 1u << b;
          tcgen05_fence_after();
          uint32_t accA[32], accB[32];
          tmem_ld(tm_lane + 128 * b + 64 * g, accA);
          tmem_ld(tm_lane + 128 * b + 64 * g + 32, accB);
          if (q == 0) {
            // every warp is past the previous layer (its last arrival gated this accumulator): stage the next bias
            const int ln = (l + 1) & (NH - 1);
            for (int i = et; i < D; i += N_EPI_W * 32) bias_s[(ln & 1) * D + i] = __ldg(bias_all + ln * D + i);
          }
          auto st32 = [&](int qq, const uint32_t *v) {   // 64 features = 32 operand columns
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { lo[k] = v[k]; hi[k] = v[16 + k]; }
            tmem_st16(tm_lane + 256 + 64 * qq + 32 * g, lo);
            tmem_st16(tm_lane + 256 + 64 * qq + 32 * g + 16, hi);
          };
          tmem_ld_wait(accA);
          tmem_ld_wait(accB);
          tcgen05_fence_before();
          arrive(a_dfree[b]);                         // the issuer may accumulate quarter q + 2 into this buffer
          const int f0 = 128 * q + 64 * g;
          uint32_t pk[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t(&a)[32] = half ? accB : accA;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const int f = f0 + 32 * half + i;
              const float4 bb = *reinterpret_cast<const float4 *>(bl + f);
              const float s0 = __sinf(__uint_as_float(a[i]) + bb.x), s1 = __sinf(__uint_as_float(a[i + 1]) + bb.y);
              const float s2 = __sinf(__uint_as_float(a[i + 2]) + bb.z), s3 = __sinf(__uint_as_float(a[i + 3]) + bb.w);
              if (last) {
                const float4 wa = *reinterpret_cast<const float4 *>(wout_s + f);
                const float4 wb = *reinterpret_cast<const float4 *>(wout_s + D + f);
                o0 += s0 * wa.x + s1 * wa.y + s2 * wa.z + s3 * wa.w;
                o1 += s0 * wb.x + s1 * wb.y + s2 * wb.z + s3 * wb.w;
              } else {
                pk[16 * half + i / 2] = pack_bf16x2(s0, s1);
                pk[16 * half + i / 2 + 1] = pack_bf16x2(s2, s3);
              }
            }
          }
          if (!last) {
            if (q < 3) {
#pragma unroll
              for (int k = 0; k < 32; ++k) held[32 * q + k] = pk[k];
              if (q == 2) {
                // Quarter 3 is being accumulated, k-block by k-block in ascending k.  Once its MMAs over k-block kb are
                // complete nothing of layer l reads operand columns k < 128 (kb + 1) any more: the held quarter kb is
                // stored under the remaining MMAs, and the next layer's first 24 instructions (k < 384) are handed to the
                // issuer before this layer's last one has run - they queue up behind it.
                // (layer 0 reads its operand from shared memory: nothing to wait for)
#pragma unroll
                for (int kb = 0; kb < 3; ++kb) {
                  if (l > 0) mbar_wait(bar.kfree(kb), kph);
                  st32(kb, held + 32 * kb);
                }
                if (l > 0) kph ^= 1;
                tmem_st_wait();
                tcgen05_fence_before();
                arrive(a_ready[1]);
              }
            } else {
              st32(3, pk);
              tmem_st_wait();
              tcgen05_fence_before();
              arrive(a_ready[2]);                     // k >= 384
            }
          }
        }
        if (last) {
          if (g > 0) osum_s[row] = make_float2(o0, o1);
          named_bar_sync(1, N_EPI_W * 32);
          if (g == 0 && m < p.M) {
            const float2 o = osum_s[row];
            p.out[m] = make_float2(o0 + o.x + __ldg(b_out) + p.off0, o1 + o.y + __ldg(b_out + 1) + p.off1);
          }
          named_bar_sync(1, N_EPI_W * 32);            // osum_s is reused by the next tile
        }
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

}  // namespace bf
}  // namespace snf

using namespace snf;

// called by snf_mlp_pack_bf16 (snf_mlp_bf16.cu)
int snf_bf16_pack_ts(const float *const *W, void *packed, cudaStream_t st) {
  const int64_t chunks = (int64_t)bf::FWD_BLOCKS * (bf::WBLK_BYTES / 16);
  bf::pack_weights_ts_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(
      W[0], W[1], W[2], W[3], W[4], W[5], W[6], W[7],
      reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_TS_OFF));
  count_launch(1);
  return launch_status();
}

// called by snf_mlp_fwd_bf16 for train == 0
int snf_bf16_forward_ts(const float *x, int64_t M, const void *packed, float off0, float off1, float *out, int num_sms,
                        cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bf::mlp_fwd_ts_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::ts::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  bf::FwdTsParams p{};
  p.x = reinterpret_cast<const float4 *>(x);
  p.M = M;
  p.num_tiles = (int)((M + bf::TILE_M - 1) / bf::TILE_M);
  p.num_tiles = (p.num_tiles + 1) / 2 * 2;
  p.packed = reinterpret_cast<const uint8_t *>(packed);
  p.out = reinterpret_cast<float2 *>(out);
  p.off0 = off0; p.off1 = off1;
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  grid &= ~1;
  bf::mlp_fwd_ts_bf16_kernel<<<grid, bf::ts::THREADS, bf::ts::SMEM_BYTES, st>>>(p);
  count_launch();
  return launch_status();
}
