// bf16 tensor-core field network for sm_100a: positional encoding + 8 x (Linear(512)+sin) + Linear(2) fused in
// ONE persistent, warp-specialised kernel per pass.  Replaces PositionalEncoding.forward / NeRF.forward /
// NeRF_DT.forward (sunerf/model/model.py:123-132, 44-57, 169-187) in "bf16-MLP mode" (BASELINE.json: 1e-2).
//
// Per CTA: one tile of 128 points at a time; the activations never leave the SM between layers:
//   warp 0      TMA producer : streams the pre-packed bf16 weights (UMMA K-major SWIZZLE_128B image) from L2 with
//                              32 KB cp.async.bulk copies through a 3-stage mbarrier ring; the fp32 W_out rows ride
//                              through the same ring as a 4 KB pseudo-block at the end of each tile
//   warp 1      MMA issuer   : one thread issues tcgen05.mma (M=128, N=256, K=16) per k-step; A = activations in
//                              shared memory (128 KB, same swizzled image), D = 128x512 fp32 in TMEM (all 512 cols)
//   warps 2-9   epilogue     : thread = (row, column half).  Double-buffered tcgen05.ld of the accumulator, + bias
//                              (staged in shared memory once per layer), sin, bf16, st.shared back into the A image
//                              for the next layer; layer 0's A image is the sin/cos encoding computed in place;
//                              the 512->2 output layer is a register dot product fused into the last epilogue
// Training mode additionally writes, per layer, the activation image h = sin(pre) (TMA bulk store straight from the
// A image) and the pre-activation image (bf16, direct 16 B stores) in the same tile-image layout for the backward.
#include "snf_bf16_common.cuh"

namespace snf {
namespace bf {

// ------------------------------------------------------------------------------------------ weight packing
// one thread per 16-byte chunk of the packed image; 32 KB blocks of 256 output features x 64 k in consumption order
// (layer, n-half, k-slab); CTA r of a pair streams rows [128r, 128r+128) of every block
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *w0, const float *w1, const float *w2,
                                                           const float *w3, const float *w4, const float *w5,
                                                           const float *w6, const float *w7, uint4 *__restrict__ dst) {
  const float *W[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)FWD_BLOCKS * (WBLK_BYTES / 16)) return;
  const int blk = (int)(idx / (WBLK_BYTES / 16));
  const int within = (int)(idx % (WBLK_BYTES / 16));
  int l, q, ks;
  if (blk < 4) { l = 0; q = blk >> 1; ks = blk & 1; }
  else { const int b2 = blk - 4; l = 1 + b2 / 16; q = (b2 % 16) >> 3; ks = b2 & 7; }
  const int r = within >> 3;                 // row inside the block (0..255)
  const int pos = within & 7;                // stored chunk position inside the 128 B row
  const int c8 = pos ^ (r & 7);              // logical chunk (SWIZZLE_128B)
  const int n = q * NCHUNK + r;
  const int kbase = ks * 64 + c8 * 8;
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
  // (r>>3)*1024 + (r&7)*128 + pos*16 == r*128 + pos*16: the image is row-linear, only the chunk order is permuted
  dst[(int64_t)blk * (WBLK_BYTES / 16) + within] = o;
}

__global__ void __launch_bounds__(256) pack_small_kernel(const float *b0, const float *b1, const float *b2,
                                                         const float *b3, const float *b4, const float *b5,
                                                         const float *b6, const float *b7, const float *w_out,
                                                         const float *b_out, float *__restrict__ dst) {
  const float *B[8] = {b0, b1, b2, b3, b4, b5, b6, b7};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NH * D) dst[i] = B[i / D][i % D];
  else if (i < NH * D + 2 * D) dst[i] = w_out[i - NH * D];
  else if (i < NH * D + 2 * D + 2) dst[i] = b_out[i - NH * D - 2 * D];
}

// ------------------------------------------------------------------------------------------ fused forward
struct FwdParams {
  const float4 *x;        // [M] (x,y,z,t)
  int64_t M;
  int num_tiles;          // rounded up to even: a CTA pair always runs two tiles
  const uint8_t *packed;  // PACK_TOTAL_BYTES
  float2 *out;            // [M]
  float off0, off1;
  uint8_t *save_enc;      // train: [tiles][2 slabs][16 KB] bf16 image, else null
  uint8_t *save_h;        // train: [tiles][8][128 KB]   sin(pre)
  uint8_t *save_pre;      // train: [tiles][8][128 KB]   pre-activation (the backward takes cos of it)
};

constexpr int FWD_RING_PER_TILE = FWD_BLOCKS + 1;   // + the W_out pseudo-block

template <bool TRAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) mlp_fwd_bf16_kernel(const FwdParams p) {
  constexpr int NSTAGE = TRAIN ? NSTAGE_TRAIN : NSTAGE_INFER;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();   // the UMMA/TMA images need a 1024-byte aligned window
  uint8_t *gA = smem_raw;
  const uint32_t sA = base, sW = base + OFF_RING;
  float *bias_s = reinterpret_cast<float *>(smem_raw + OFF_BIAS);
  const Bars bar{base + OFF_BAR};
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // leader: full[s] = its own TMA + one remote arrival from the peer's relay
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar.full(s), rank == 0 ? 2 : 1); mbar_init(bar.empty(s), 1); }
    mbar_init(bar.acc(), 1);
    mbar_init(bar.aready(), 2);        // leader: its own epilogue warps + the peer's
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(bar.tmem_slot(), 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem_raw + TMEM_SLOT_OFF);

  const float *bias_all = reinterpret_cast<const float *>(p.packed + PACK_BIAS_OFF);
  const float *b_out = reinterpret_cast<const float *>(p.packed + PACK_BOUT_OFF);

  if (warp == 0) {
    // =========================== TMA producer (each CTA streams its half of every weight block)
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        for (int blk = 0; blk < FWD_RING_PER_TILE; ++blk) {
          mbar_wait_cluster(bar.empty(s), ph ^ 1);
          if (blk < FWD_BLOCKS) {
            mbar_arrive_expect_tx(bar.full(s), WHALF_BYTES);
            bulk_g2s(sW + s * WHALF_BYTES, p.packed + (int64_t)blk * WBLK_BYTES + rank * WHALF_BYTES, WHALF_BYTES, bar.full(s));
          } else {
            mbar_arrive_expect_tx(bar.full(s), WOUT_BYTES);
            bulk_g2s(sW + s * WHALF_BYTES, p.packed + PACK_WOUT_OFF, WOUT_BYTES, bar.full(s));
          }
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0, ph_a = 0;
      if (rank == 0) {
        // =========================== MMA issuer (leader CTA): M=256 across the pair, N=256 per instruction
        const uint32_t idesc = idesc_bf16(256, NCHUNK);
        for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
          for (int l = 0; l < NH; ++l) {
            mbar_wait_cluster(bar.aready(), ph_a); ph_a ^= 1;   // both A images complete, both TMEMs drained
            tcgen05_fence_after();
            const int nslab = l == 0 ? 2 : 8;
            for (int q = 0; q < 2; ++q) {
              for (int ks = 0; ks < nslab; ++ks) {
                mbar_wait_cluster(bar.full(s), ph);         // both halves of the stage have landed
                tcgen05_fence_after();
                const int ksteps = (l == 0 && ks == 1) ? (K0 - 64) / 16 : 4;
#pragma unroll 4
                for (int k4 = 0; k4 < ksteps; ++k4) {
                  const uint64_t ad = smem_desc(sA + ks * SLAB_BYTES + k4 * 32, 16, 1024);
                  const uint64_t bd = smem_desc(sW + s * WHALF_BYTES + k4 * 32, 16, 1024);
                  mma_ss_2cta(tmem + q * NCHUNK, ad, bd, idesc, (ks | k4) != 0);
                }
                mma_commit_2cta(bar.empty(s), 3);           // frees the stage in both CTAs
                if (++s == NSTAGE) { s = 0; ph ^= 1; }
              }
            }
            mma_commit_2cta(bar.acc(), 3);                  // whole layer accumulated, in both CTAs
          }
          if (++s == NSTAGE) { s = 0; ph ^= 1; }            // the W_out pseudo-block is consumed by the epilogue warps
        }
      } else {
        // =========================== peer relay: tell the leader when this CTA's half of a stage has landed
        // every ring slot is relayed (the W_out pseudo-block too) so the leader's full[s] keeps the ring's phase
        for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
          for (int blk = 0; blk < FWD_RING_PER_TILE; ++blk) {
            mbar_wait(bar.full(s), ph);
            mbar_arrive_remote(mapa_shared(bar.full(s), 0));
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // =========================== epilogue warps: 256 threads, thread = (row, column half)
    const int e = warp - 2;
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int half = e >> 2;                          // columns [256*half, 256*half + 256)
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;                  // 0..255
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16) + half * 256;
    constexpr bool train = TRAIN;
    uint8_t *stg = smem_raw + OFF_STG + e * STG_WARP_BYTES;       // this warp's 4 KB staging block
    const uint32_t stg_s = base + OFF_STG + e * STG_WARP_BYTES;
    uint32_t ph_acc = 0;
    int ring_pos = 0;                                 // ring slots consumed by earlier tiles
    auto signal_aready = [&]() {                      // all 256 threads: fences done -> one arrival per CTA
      tcgen05_fence_before();
      named_bar_sync(1, N_EPI);
      if (et == 0) {
        if (rank == 0) mbar_arrive(bar.aready());
        else mbar_arrive_remote(mapa_shared(bar.aready(), 0));
      }
    };
    // chunks 4..7 of slab 1 are never read by the layer-0 MMAs but are part of the saved encoder image
    if (half == 1)
      for (int c8 = 4; c8 < 8; ++c8)
        *reinterpret_cast<uint4 *>(gA + SLAB_BYTES + sw128_chunk_off(row, c8)) = make_uint4(0, 0, 0, 0);
    for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs, ring_pos += FWD_RING_PER_TILE) {
      const int tile = tp * 2 + (int)rank;
      const int64_t m = (int64_t)tile * TILE_M + row;
      // ---- layer-0 operand: positional encoding of this row, written straight into the A image.
      //      half 0: x and frequencies 0..4 ; half 1: bf16 residual of x, zero padding and frequencies 5..9
      {
        if (train) { if (et == 0) bulk_wait_read_all(); named_bar_sync(1, N_EPI); }
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M) xv = p.x[m];
        const float xc[4] = {xv.x, xv.y, xv.z, xv.w};
        auto put8 = [&](int feat0, float a, float b, float c, float d) {   // 4 consecutive features (8 bytes)
          const int c8 = feat0 >> 3;
          uint2 v = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
          *reinterpret_cast<uint2 *>(gA + (c8 >> 3) * SLAB_BYTES + sw128_chunk_off(row, c8 & 7) + (feat0 & 7) * 2) = v;
        };
        if (half == 0) {
          put8(0, xc[0], xc[1], xc[2], xc[3]);
        } else {
          float rs[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) rs[c] = xc[c] - __bfloat162float(__float2bfloat16_rn(xc[c]));   // what bf16 drops from x
          put8(84, rs[0], rs[1], rs[2], rs[3]);
          put8(88, 0.f, 0.f, 0.f, 0.f);
          put8(92, 0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int fi = 0; fi < 5; ++fi) {
          const int f = half * 5 + fi;
          float sv[4], cv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) sincosf(xc[c] * (float)(1 << f) * 0.5f, &sv[c], &cv[c]);   // exact scalings
          put8(4 + f * 4, sv[0], sv[1], sv[2], sv[3]);
          put8(44 + f * 4, cv[0], cv[1], cv[2], cv[3]);
        }
        fence_proxy_async_smem();
        if (train) {
          named_bar_sync(1, N_EPI);
          if (et == 0) {
            bulk_s2g(p.save_enc + (int64_t)tile * 2 * SLAB_BYTES, sA, 2 * SLAB_BYTES);
            bulk_commit();
          }
        }
        signal_aready();
      }
      // ---- layers
      for (int l = 0; l < NH; ++l) {
        const float2 bnext = __ldg(reinterpret_cast<const float2 *>(bias_all + l * D) + et);   // in flight during the MMAs
        mbar_wait_cluster(bar.acc(), ph_acc); ph_acc ^= 1;
        tcgen05_fence_after();
        const bool last = (l == NH - 1);
        const bool write_a = !last || train;
        if (train && et == 0) bulk_wait_read_all();   // the previous bulk store has finished reading the A image
        reinterpret_cast<float2 *>(bias_s)[et] = bnext;
        named_bar_sync(1, N_EPI);
        const float *wout_s = nullptr;
        if (last) {   // W_out pseudo-block: ring slot ring_pos + FWD_BLOCKS
          const int slot = ring_pos + FWD_BLOCKS;
          mbar_wait_cluster(bar.full(slot % NSTAGE), (uint32_t)((slot / NSTAGE) & 1));
          wout_s = reinterpret_cast<const float *>(smem_raw + OFF_RING + (slot % NSTAGE) * WHALF_BYTES);
        }
        uint8_t *psave = train ? p.save_pre + ((int64_t)tile * NH + l) * A_BYTES : nullptr;
        float o0 = 0.f, o1 = 0.f;
        uint32_t accA[32], accB[32];
        tmem_ld32(tm_row, accA);
        auto process = [&](const uint32_t (&acc)[32], int g) {
          const int col0 = half * 256 + g * 32;
          float hv[32];
          uint32_t ppk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4 *>(bias_s + col0 + i);
            const float v0 = __uint_as_float(acc[i]) + b.x, v1 = __uint_as_float(acc[i + 1]) + b.y;
            const float v2 = __uint_as_float(acc[i + 2]) + b.z, v3 = __uint_as_float(acc[i + 3]) + b.w;
            hv[i] = __sinf(v0); hv[i + 1] = __sinf(v1); hv[i + 2] = __sinf(v2); hv[i + 3] = __sinf(v3);
            if (train) { ppk[i / 2] = pack_bf16x2(v0, v1); ppk[i / 2 + 1] = pack_bf16x2(v2, v3); }
          }
          if (last) {   // fused output layer: out = W_out h + b_out
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 wa = *reinterpret_cast<const float4 *>(wout_s + col0 + i);
              const float4 wb = *reinterpret_cast<const float4 *>(wout_s + D + col0 + i);
              o0 += hv[i] * wa.x + hv[i + 1] * wa.y + hv[i + 2] * wa.z + hv[i + 3] * wa.w;
              o1 += hv[i] * wb.x + hv[i + 1] * wb.y + hv[i + 2] * wb.z + hv[i + 3] * wb.w;
            }
          }
          const int slab = col0 >> 6;
          if (train && (g & 1) == 0) {   // new slab: the previous staged block must have left shared memory
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int c8 = ((col0 & 63) >> 3) + c;
            if (write_a) {
              uint4 o;
              o.x = pack_bf16x2(hv[c * 8 + 0], hv[c * 8 + 1]); o.y = pack_bf16x2(hv[c * 8 + 2], hv[c * 8 + 3]);
              o.z = pack_bf16x2(hv[c * 8 + 4], hv[c * 8 + 5]); o.w = pack_bf16x2(hv[c * 8 + 6], hv[c * 8 + 7]);
              *reinterpret_cast<uint4 *>(gA + slab * SLAB_BYTES + sw128_chunk_off(row, c8)) = o;
            }
            if (train)   // pre-activation: staged per warp (32 rows x 128 B, image layout), then one 4 KB bulk store
              *reinterpret_cast<uint4 *>(stg + sw128_chunk_off(lane, c8)) = make_uint4(ppk[c * 4], ppk[c * 4 + 1], ppk[c * 4 + 2], ppk[c * 4 + 3]);
          }
          if (train && (g & 1) == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              bulk_s2g(psave + slab * SLAB_BYTES + (q * 32) * 128, stg_s, STG_WARP_BYTES);
              bulk_commit();
            }
          }
        };
#pragma unroll 1
        for (int g = 0; g < 8; g += 2) {   // software-pipelined TMEM reads: the next group loads while this one computes
          tmem_ld_wait(accA);
          tmem_ld32(tm_row + (g + 1) * 32, accB);
          process(accA, g);
          tmem_ld_wait(accB);
          if (g + 2 < 8) tmem_ld32(tm_row + (g + 2) * 32, accA);
          process(accB, g + 1);
        }
        if (write_a) fence_proxy_async_smem();
        if (train) {
          named_bar_sync(1, N_EPI);
          if (et == 0) {
            uint8_t *dst = p.save_h + ((int64_t)tile * NH + l) * A_BYTES;
#pragma unroll 1
            for (int sl = 0; sl < 8; ++sl) bulk_s2g(dst + sl * SLAB_BYTES, sA + sl * SLAB_BYTES, SLAB_BYTES);
            bulk_commit();
          }
        }
        if (last) {
          // combine the two column halves of each row through the (now idle) bias buffer, release the W_out slot
          named_bar_sync(1, N_EPI);
          if (et == 0) mbar_arrive(bar.empty((ring_pos + FWD_BLOCKS) % NSTAGE));
          if (half == 1) reinterpret_cast<float2 *>(bias_s)[row] = make_float2(o0, o1);
          named_bar_sync(1, N_EPI);
          if (half == 0 && m < p.M) {
            const float2 o = reinterpret_cast<const float2 *>(bias_s)[row];
            p.out[m] = make_float2((o0 + o.x) + __ldg(b_out) + p.off0, (o1 + o.y) + __ldg(b_out + 1) + p.off1);
          }
        } else {
          signal_aready();
        }
      }
    }
    if (train && lane == 0) bulk_wait_all();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

}  // namespace bf
}  // namespace snf

using namespace snf;

// snf_mlp_bf16_bwd.cu
int snf_bf16_pack_wt(const float *const *W, void *packed, cudaStream_t st);
int snf_bf16_backward(const float *grad_out, int64_t M, const void *packed, const bf::Bf16Ws &w, float *const *gW,
                      float *const *gB, int num_sms, cudaStream_t st);

int64_t snf_mlp_bf16_ws_bytes(int64_t M, int train) { return bf::bf16_layout(nullptr, M, train).bytes; }

extern "C" int64_t snf_mlp_pack_bytes(void) { return bf::PACK_TOTAL_BYTES; }

extern "C" int snf_mlp_pack_bf16(const float *const *W, const float *const *B, void *packed, void *stream) {
  SNF_CHECK_PTR(W); SNF_CHECK_PTR(B); SNF_CHECK_PTR(packed); SNF_CHECK_ALIGN(packed, 1024);
  for (int l = 0; l <= bf::NH; ++l) { SNF_CHECK_PTR(W[l]); SNF_CHECK_PTR(B[l]); }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = (int64_t)bf::FWD_BLOCKS * (bf::WBLK_BYTES / 16);
  bf::pack_weights_kernel<<<(unsigned)ceil_div64(chunks, 256), 256, 0, st>>>(W[0], W[1], W[2], W[3], W[4], W[5], W[6], W[7],
                                                                          reinterpret_cast<uint4 *>(packed));
  const int nsmall = bf::NH * bf::D + 2 * bf::D + 2;
  bf::pack_small_kernel<<<(nsmall + 255) / 256, 256, 0, st>>>(B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], W[8], B[8],
                                                             reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_BIAS_OFF));
  count_launch(2);
  return snf_bf16_pack_wt(W, packed, st);   // W^T blocks for the dgrad chain
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms = n > 0 ? n : 148;
  }
  return g_num_sms;
}

extern "C" int snf_mlp_fwd_bf16(const float *x, int64_t M, const void *packed, float off0, float off1, float *out,
                                void *ws, int train, void *stream) {
  SNF_CHECK_PTR(x); SNF_CHECK_PTR(packed); SNF_CHECK_PTR(out);
  SNF_CHECK_ALIGN(x, 16); SNF_CHECK_ALIGN(out, 8); SNF_CHECK_ALIGN(packed, 1024);
  if (M < 0) return SNF_E_ARG;
  if (M == 0) return 0;
  if (train) { SNF_CHECK_PTR(ws); SNF_CHECK_ALIGN(ws, 1024); }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bf::mlp_fwd_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(bf::mlp_fwd_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  bf::FwdParams p{};
  p.x = reinterpret_cast<const float4 *>(x);
  p.M = M;
  p.num_tiles = (int)((M + bf::TILE_M - 1) / bf::TILE_M);
  p.num_tiles = (p.num_tiles + 1) / 2 * 2;
  p.packed = reinterpret_cast<const uint8_t *>(packed);
  p.out = reinterpret_cast<float2 *>(out);
  p.off0 = off0; p.off1 = off1;
  if (train) {
    bf::Bf16Ws w = bf::bf16_layout(ws, M, 1);
    p.save_enc = w.enc; p.save_h = w.h; p.save_pre = w.pre;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  grid &= ~1;   // whole CTA pairs
  if (train) bf::mlp_fwd_bf16_kernel<true><<<grid, bf::NTHREADS, bf::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  else bf::mlp_fwd_bf16_kernel<false><<<grid, bf::NTHREADS, bf::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  count_launch();
  return launch_status();
}

extern "C" int snf_mlp_bwd_bf16(const float *x, int64_t M, const void *packed, const float *grad_out, void *ws,
                                float *const *gW, float *const *gB, void *stream) {
  (void)x;
  SNF_CHECK_PTR(packed); SNF_CHECK_PTR(grad_out); SNF_CHECK_PTR(ws); SNF_CHECK_PTR(gW); SNF_CHECK_PTR(gB);
  SNF_CHECK_ALIGN(grad_out, 8); SNF_CHECK_ALIGN(ws, 1024); SNF_CHECK_ALIGN(packed, 1024);
  if (M <= 0) return SNF_E_ARG;
  for (int l = 0; l <= bf::NH; ++l) { SNF_CHECK_PTR(gW[l]); SNF_CHECK_PTR(gB[l]); SNF_CHECK_ALIGN(gW[l], 16); }
  bf::Bf16Ws w = bf::bf16_layout(ws, M, 1);
  return snf_bf16_backward(grad_out, M, packed, w, gW, gB, num_sms(), (cudaStream_t)stream);
}
