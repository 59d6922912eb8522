// 16-bit tensor-core field network for sm_100a: positional encoding + 8 x (Linear(512)+sin) + Linear(2) fused in
// ONE persistent, warp-specialised kernel per pass.  Replaces PositionalEncoding.forward / NeRF.forward /
// NeRF_DT.forward (sunerf/model/model.py:123-132, 44-57, 169-187) in "bf16-MLP mode" (BASELINE.json: 1e-2).
// Operands are fp16 (weights and activations; 11-bit significands, see snf_bf16_common.cuh), accumulation fp32.
//
// One persistent CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 points, 640 threads per CTA; the activations never
// leave the SM between layers:
//   warp 0      TMA producer : streams this CTA's half (128 output features, 16 KB) of every pre-packed fp16 weight
//                              block (UMMA K-major SWIZZLE_128B image, L2 evict_last) through a 5-stage mbarrier ring
//   warp 1      MMA issuer   : leader CTA: tcgen05.mma M=256 (pair) x N=256 x K=16, A = activation image in shared
//                              memory (128 KB), D = 128 x 512 fp32 = all of TMEM; peer CTA: relays "my half of the
//                              stage has landed" to the leader with a relaxed remote mbarrier arrive
//   warp 2      store warp   : (training) TMA-stores every finished slab of the A image for the backward
//   warps 4-19  epilogue     : thread = (row, 16-column group).  A layer is accumulated as two temporal N-halves; the
//                              epilogue of half 0 (TMEM -> +bias -> sin -> fp16) runs under the MMAs of half 1, keeps its
//                              result in registers and writes it half-way through half 1 (acc1a: no MMA of the layer
//                              reads slabs 0-3 any more); half 1 is written slab by slab, so the next layer's MMAs queue
//                              up behind this layer's last one and continue k-slab by k-slab.
//                              Layer 0's A image is the sin/cos encoding computed in place; the 512->2 output layer
//                              is a register dot product fused into the last epilogue.
//   setmaxnreg  40 registers for the control warpgroup, 104 for the four epilogue warpgroups.
// Training mode additionally writes, per layer, the activation image h = sin(pre) (16 KB TMA bulk stores straight from the
// A image, issued by the store warp) and cos(pre) as one-byte codes (coalesced 16 B stores, chunk-major layout) for the backward.
// 16 epilogue warps (4 per SM sub-partition) hide the MUFU / TMEM-load latencies of the sine epilogue better than 8:
// measured -7 % on the forward; the dgrad epilogue (no MUFU) is faster with 8 warps and more registers.
#define SNF_EPI_GROUPS 4
#include "snf_bf16_common.cuh"

namespace snf {
namespace bf {

// ------------------------------------------------------------------------------------------ weight packing
// one thread per 16-byte chunk of the packed image; 32 KB blocks of 256 output features x 64 k in consumption order
// (layer, n-half, k-slab); CTA r of a pair streams rows [128r, 128r+128) of every block
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *w0, const float *w1, const float *w2,
                                                           const float *w3, const float *w4, const float *w5,
                                                           const float *w6, const float *w7, uint4 *__restrict__ dst,
                                                           uint4 *__restrict__ dst_lo) {
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
  o.x = pack_f16x2(v[0], v[1]); o.y = pack_f16x2(v[2], v[3]);
  o.z = pack_f16x2(v[4], v[5]); o.w = pack_f16x2(v[6], v[7]);
  // (r>>3)*1024 + (r&7)*128 + pos*16 == r*128 + pos*16: the image is row-linear, only the chunk order is permuted
  dst[(int64_t)blk * (WBLK_BYTES / 16) + within] = o;
  // low halves fp16(W - fp16(W)) for the split-precision forward (snf_mlp_x3.cu)
  float lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lo[i] = v[i] - __half2float(__float2half_rn(v[i]));
  o.x = pack_f16x2(lo[0], lo[1]); o.y = pack_f16x2(lo[2], lo[3]);
  o.z = pack_f16x2(lo[4], lo[5]); o.w = pack_f16x2(lo[6], lo[7]);
  dst_lo[(int64_t)blk * (WBLK_BYTES / 16) + within] = o;
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
  uint8_t *save_enc;      // train: [tiles][2 slabs][16 KB] fp16 image, else null
  uint8_t *save_h;        // train: [tiles][8][128 KB]   sin(pre)
  uint8_t *save_pre;      // train: [tiles][8][64 KB]    cos(pre) codes (C_BYTES layout, cosq_enc) for the dgrad chain
};

#ifndef SNF_TRACE_TRAIN
#define SNF_TRACE_TRAIN 0
#endif
#ifdef SNF_PROF
__device__ unsigned long long g_prof[2][148 * 8];
__device__ long long g_trace[4][512];   // CTA 0, inference: clock stamps per ring block
#define PROF_DECL(n) long long n = 0
#define PROF_T0(t) const long long t = clock64()
#define PROF_ADD(n, t) n += clock64() - t
#else
#define PROF_DECL(n)
#define PROF_T0(t)
#define PROF_ADD(n, t)
#endif

__device__ __forceinline__ void st_stream16(uint8_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  __stcs(reinterpret_cast<uint4 *>(dst), make_uint4(a, b, c, d));   // written once, read by the backward much later
}

template <bool TRAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) mlp_fwd_bf16_kernel(const FwdParams p) {
  constexpr int NSTAGE = fw::NSTAGE, OFF_RING = fw::OFF_RING, OFF_BIAS = fw::OFF_BIAS, OFF_WOUT = fw::OFF_WOUT,
                OFF_OSUM = fw::OFF_OSUM, OFF_BAR = fw::OFF_BAR;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();   // the UMMA/TMA images need a 1024-byte aligned window
  uint8_t *gA = smem_raw;
  const uint32_t sA = base, sW = base + OFF_RING;
  float *bias_s = reinterpret_cast<float *>(smem_raw + OFF_BIAS);          // [2][512]: layer l lives in buffer l & 1
  float *wout_s = reinterpret_cast<float *>(smem_raw + OFF_WOUT);          // [2][512]
  float2 *osum_s = reinterpret_cast<float2 *>(smem_raw + OFF_OSUM);        // [128]
  const fw::Bars bar{base + OFF_BAR};
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *bias_all = reinterpret_cast<const float *>(p.packed + PACK_BIAS_OFF);
  const float *b_out = reinterpret_cast<const float *>(p.packed + PACK_BOUT_OFF);

  if (threadIdx.x == 0) {
    // leader: full[s] = its own TMA + one remote arrival from the peer's relay
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar.full(s), rank == 0 ? 2 : 1); mbar_init(bar.empty(s), 1); }
    mbar_init(bar.acc(0), 1); mbar_init(bar.acc(1), 1);
    for (int k = 0; k < 5; ++k) mbar_init(bar.ready(k), 2 * N_EPI_WARPS);
    for (int k = 0; k < 8; ++k) mbar_init(bar.wrote(k), N_EPI_WARPS);
    mbar_init(bar.afree(), 1);
    mbar_init(bar.acc1a(), 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * D; i += NTHREADS) wout_s[i] = __ldg(reinterpret_cast<const float *>(p.packed + PACK_WOUT_OFF) + i);
  // training: the epilogue works on pre / 2 (half-angle pair, see below), so the staged bias is b / 2
  constexpr float BSC = TRAIN ? 0.5f : 1.f;
  for (int i = threadIdx.x; i < D; i += NTHREADS) bias_s[i] = BSC * __ldg(bias_all + i);
  if (warp == 1) tmem_alloc_2cta(bar.tmem_slot(), 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem_raw + fw::TMEM_SLOT_OFF);

  if (warp < EPI_WARP0) {
  reg_dealloc<REGS_CTRL>();
  if (warp == 0) {
    // =========================== TMA producer (each CTA streams its half of every weight block)
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const uint64_t keep = l2_policy_evict_last();
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        for (int blk = 0; blk < FWD_BLOCKS; ++blk) {
          mbar_wait(bar.empty(s), ph ^ 1);
#ifdef SNF_PROF
          if (TRAIN == SNF_TRACE_TRAIN && blockIdx.x == 0 && tp == pair + npairs && blk < 512) g_trace[0][blk] = clock64();
#endif
          mbar_arrive_expect_tx(bar.full(s), WHALF_BYTES);
          bulk_g2s_hint(sW + s * WHALF_BYTES, p.packed + (int64_t)blk * WBLK_BYTES + rank * WHALF_BYTES, WHALF_BYTES, bar.full(s), keep);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    if (rank == 0) {
      // =========================== MMA issuer (leader CTA): M=256 across the pair, N=256 per instruction.
      // Layer l is accumulated as two temporal halves (D columns [0,256) then [256,512)); the first half of layer
      // l+1 starts as soon as the epilogue has produced its k-slabs, while it is still draining half 1 of layer l.
      // The whole warp runs the (uniform) control flow, one elected lane issues: descriptors stay in uniform registers.
      const uint32_t idesc = idesc_f16kind(256, NCHUNK, FMT, FMT);
      const uint64_t adesc0 = smem_desc(sA, 16, 1024), bdesc0 = smem_desc(sW, 16, 1024);
      uint32_t rph = 0;
      PROF_DECL(t_ready); PROF_DECL(t_full); PROF_T0(t_begin);
      auto wait_ready = [&](int k) {
        PROF_T0(t0);
        mbar_wait(bar.ready(k), (rph >> k) & 1u); rph ^= 1u << k;
        PROF_ADD(t_ready, t0);
      };
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        int tblk = 0;
        for (int l = 0; l < NH; ++l) {
          const int nslab = l == 0 ? 2 : 8;
          for (int h = 0; h < 2; ++h) {
            if (l == 0 && h == 1) wait_ready(4);          // D half 1 drained by the previous tile's last epilogue
            for (int ks = 0; ks < nslab; ++ks) {
              if (h == 0) {
                if (ks == 0) wait_ready(0);               // slabs 0..3 of the new A image (and D half 0 drained)
                else if (ks >= 4) wait_ready(ks - 3);     // slab ks (ks == 7: D half 1 drained as well)
              }
#ifdef SNF_PROF
              if (TRAIN == SNF_TRACE_TRAIN && blockIdx.x == 0 && tp == pair + npairs && lane == 0) g_trace[1][tblk] = clock64();
#endif
              {
                PROF_T0(t0);
                mbar_wait(bar.full(s), ph);               // both halves of the stage have landed
                PROF_ADD(t_full, t0);
              }
#ifdef SNF_PROF
              if (TRAIN == SNF_TRACE_TRAIN && blockIdx.x == 0 && tp == pair + npairs && lane == 0) g_trace[2][tblk] = clock64();
              ++tblk;
#endif
              tcgen05_fence_after();
              if (elect_one()) {
                const uint64_t ad = adesc0 + (uint64_t)((ks * SLAB_BYTES) >> 4), bd = bdesc0 + (uint64_t)((s * WHALF_BYTES) >> 4);
                if (l == 0 && ks == 1) {
#pragma unroll
                  for (int k4 = 0; k4 < (K0 - 64) / 16; ++k4) mma_ss_2cta(tmem + h * NCHUNK, ad + 2 * k4, bd + 2 * k4, idesc, 1);
                } else {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) mma_ss_2cta(tmem + h * NCHUNK, ad + 2 * k4, bd + 2 * k4, idesc, (ks | k4) != 0);
                }
                mma_commit_2cta(bar.empty(s), 3);         // frees the stage in both CTAs
                if (ks == nslab - 1) mma_commit_2cta(bar.acc(h), 3);   // this half of the layer is accumulated, in both CTAs
                if (h == 1 && ks == (nslab < 4 ? nslab : 4) - 1) mma_commit_2cta(bar.acc1a(), 3);   // slabs 0..3 of A are no longer read
              }
              __syncwarp();
              if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
          }
        }
      }
#ifdef SNF_PROF
      if (lane == 0) {
        g_prof[TRAIN][blockIdx.x * 8 + 0] = clock64() - t_begin;
        g_prof[TRAIN][blockIdx.x * 8 + 1] = t_ready;
        g_prof[TRAIN][blockIdx.x * 8 + 2] = t_full;
      }
#endif
    } else if (lane == 0) {
      // =========================== peer relay: tell the leader when this CTA's half of a stage has landed
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        for (int blk = 0; blk < FWD_BLOCKS; ++blk) {
          mbar_wait(bar.full(s), ph);
          mbar_arrive_remote_relaxed(mapa_shared(bar.full(s), 0));   // the data is TMA-written and tensor-core-read
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (TRAIN && warp == 2 && lane == 0) {
    // =========================== store warp (training): TMA-stores every finished slab of the A image (the encoder
    // image, then h_l of each layer) for the backward.  The epilogue warps only arrive on wrote[] (non-blocking): the
    // issue latency of the bulk stores - they back up behind the HBM writes - stays off the epilogue's critical path.
    const uint64_t stream_pol = l2_policy_evict_first();   // written once, read by the backward much later
    uint32_t wph = 0;
    auto wait_wrote = [&](int slot) { mbar_wait(bar.wrote(slot), (wph >> slot) & 1u); wph ^= 1u << slot; };
    auto store_slabs = [&](uint8_t *img, int sl0, int nsl) {
      for (int sl = sl0; sl < sl0 + nsl; ++sl) bulk_s2g_hint(img + sl * SLAB_BYTES, sA + sl * SLAB_BYTES, SLAB_BYTES, stream_pol);
      bulk_commit();
    };
    for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
      const int tile = tp * 2 + (int)rank;
      wait_wrote(0);
      store_slabs(p.save_enc + (int64_t)tile * 2 * SLAB_BYTES, 0, 2);
      bulk_wait_read_all();
      mbar_arrive(bar.afree());                         // the encoder image may be overwritten by h_0
      for (int l = 0; l < NH; ++l) {
        uint8_t *hsave = p.save_h + ((int64_t)tile * NH + l) * A_BYTES;
        wait_wrote(0);
        store_slabs(hsave, 0, 4);
        for (int j = 0; j < 4; ++j) { wait_wrote(4 + j); store_slabs(hsave, 4 + j, 1); }
        bulk_wait_read_all();
        mbar_arrive(bar.afree());                       // h_l has left the A image
      }
    }
    bulk_wait_all();
  }
  } else {
    reg_alloc<REGS_EPI>();
    // =========================== epilogue warps.  Thread = (row, g): in temporal half h, step j it owns the CPT
    // accumulator columns 256h + 64j + CPT g .. , i.e. chunks CHUNKS g .. of A slab 4h + j.
    const int e = warp - EPI_WARP0;
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int g = e >> 2;                             // column group inside a step (0 .. EPI_GROUPS-1)
    const int row = q * 32 + lane;
    const int et = threadIdx.x - EPI_WARP0 * 32;      // 0 .. N_EPI-1
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16) + g * CPT;
    uint32_t ready_addr[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ready_addr[k] = rank == 0 ? bar.ready(k) : mapa_shared(bar.ready(k), 0);
#ifdef SNF_PROF
    int estamp = 0; bool etrace = false;
#define ESTAMP() do { if (etrace && e == 0 && lane == 0 && estamp < 512) g_trace[3][estamp++] = clock64(); } while (0)
#else
#define ESTAMP()
#endif
    // one arrival per warp: every lane has fenced its own writes, __syncwarp orders them before lane 0's arrival
    auto arrive_ready = [&](int k) {
      __syncwarp();
      ESTAMP();
      if (lane == 0) {
        // the writes were handed to the async proxy by each lane's fence.proxy.async; the arrival itself is a signal
        if (rank == 0) mbar_arrive(ready_addr[k]);
        else mbar_arrive_remote_relaxed(ready_addr[k]);
      }
    };
    // Training: after a slab (or slabs 0-3 / the encoder image) is written and fenced, the warp arrives on the CTA-local
    // wrote[slot] barrier - the store warp issues the TMA stores - and before the A image is overwritten it waits for
    // afree (the previous layer's stores have finished reading it; normally long complete).
    auto arrive_wrote = [&](int slot) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar.wrote(slot));
    };
    uint32_t aph = 0;
    bool first_tile = true;
    PROF_DECL(t_stw);
    auto wait_afree = [&]() {
      PROF_T0(t0);
      mbar_wait(bar.afree(), aph); aph ^= 1;
      PROF_ADD(t_stw, t0);
    };
    uint32_t ph = 0;
    PROF_DECL(t_acc0); PROF_DECL(t_acc1); PROF_DECL(t_enc); PROF_T0(t_begin);
    tcgen05_fence_before();
    arrive_ready(4);                                  // D half 1 is free for the first tile
    for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
      const int tile = tp * 2 + (int)rank;
      const int64_t m = (int64_t)tile * TILE_M + row;
#ifdef SNF_PROF
      etrace = TRAIN == SNF_TRACE_TRAIN && blockIdx.x == 0 && tp == pair + npairs;
#endif
      // ---- layer-0 operand: positional encoding of this row, written straight into the A image (slabs 0, 1).
      //      The 10 frequencies are dealt round-robin to the column groups; group 0 adds the raw coordinates, the last
      //      group the fp16 residual of x and the zero padding.
      {
        PROF_T0(t0);
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M) xv = p.x[m];
        const float xc[4] = {xv.x, xv.y, xv.z, xv.w};
        if (TRAIN && !first_tile) wait_afree();       // the previous tile's last layer must have left the A image
        first_tile = false;
        auto put8 = [&](int feat0, float a, float b, float c, float d) {   // 4 consecutive features (8 bytes)
          const int c8 = feat0 >> 3;
          *reinterpret_cast<uint2 *>(gA + (c8 >> 3) * SLAB_BYTES + sw128_chunk_off(row, c8 & 7) + (feat0 & 7) * 2) =
              make_uint2(pack_f16x2(a, b), pack_f16x2(c, d));
        };
        if (g == 0) put8(0, xc[0], xc[1], xc[2], xc[3]);
        if (g == EPI_GROUPS - 1) {
          float rs[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) rs[c] = xc[c] - __half2float(__float2half_rn(xc[c]));   // what fp16 drops from x
          put8(84, rs[0], rs[1], rs[2], rs[3]);
          put8(88, 0.f, 0.f, 0.f, 0.f);
          put8(92, 0.f, 0.f, 0.f, 0.f);
        }
        if (TRAIN && g == EPI_GROUPS - 2)   // features 96..127 of the saved encoder image (read by the layer-0 wgrad only)
          for (int c8 = 4; c8 < 8; ++c8)
            *reinterpret_cast<uint4 *>(gA + SLAB_BYTES + sw128_chunk_off(row, c8)) = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int fi = 0; fi < (10 + EPI_GROUPS - 1) / EPI_GROUPS; ++fi) {
          const int f = fi * EPI_GROUPS + g;
          if (f < 10) {
            float sv[4], cv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) sincosf(xc[c] * (float)(1 << f) * 0.5f, &sv[c], &cv[c]);   // exact scalings
            put8(4 + f * 4, sv[0], sv[1], sv[2], sv[3]);
            put8(44 + f * 4, cv[0], cv[1], cv[2], cv[3]);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        arrive_ready(0);
        if (TRAIN) arrive_wrote(0);                   // the store warp saves the encoder image (slabs 0, 1)
        PROF_ADD(t_enc, t0);
      }
      // ---- layers
#pragma unroll 1
      for (int l = 0; l < NH; ++l) {
        const bool last = (l == NH - 1);
        const float *bl = bias_s + (l & 1) * D;
        uint8_t *psave = TRAIN ? p.save_pre + ((int64_t)tile * NH + l) * C_BYTES : nullptr;
        uint32_t held[4 * CPT / 2];                   // half 0 of h_l (fp16 pairs): the MMAs of half 1 still read A
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          {
            PROF_T0(t0);
            mbar_wait(bar.acc(h), ph);
            if (h == 0) { PROF_ADD(t_acc0, t0); } else { PROF_ADD(t_acc1, t0); }
          }
          ESTAMP();
          tcgen05_fence_after();
          uint32_t accA[CPT], accB[CPT];
          tmem_ld(tm_row + h * 256, accA);
          if (h == 0) {
            // every warp is past the previous layer: the other bias buffer is idle -> stage the next layer's bias
            const int ln = (l + 1) & (NH - 1);
            for (int i = et; i < D; i += N_EPI) bias_s[(ln & 1) * D + i] = BSC * __ldg(bias_all + ln * D + i);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t(&cur)[CPT] = (j & 1) ? accB : accA;
            uint32_t(&nxt)[CPT] = (j & 1) ? accA : accB;
            tmem_ld_wait(cur);
            if (j + 1 < 4) tmem_ld(tm_row + h * 256 + (j + 1) * 64, nxt);
            const int col0 = h * 256 + j * 64 + g * CPT;
            const int sl = h * 4 + j;
            uint32_t pk[CPT / 2], cq[CPT / 4];
#pragma unroll
            for (int i = 0; i < CPT; i += 4) {
              const float4 b = *reinterpret_cast<const float4 *>(bl + col0 + i);
              float s0, s1, s2, s3;
              if (TRAIN) {
                // Half-angle pair per element (two MUFU, as sin + cos of the full angle would be): with s = sin(pre/2),
                // c = cos(pre/2): sin(pre) = 2 s c and the backward's cosine code comes from min(|s|, |c|) and |s| > |c|.
                // Everything after the MUFUs runs on fp16 PAIRS (the results are fp16 anyway): 9.5 instead of 14
                // instructions per element - the epilogue of a half layer must stay shorter than the MMAs of the other half.
                const float h0 = fmaf(__uint_as_float(cur[i]), 0.5f, b.x), h1 = fmaf(__uint_as_float(cur[i + 1]), 0.5f, b.y);
                const float h2 = fmaf(__uint_as_float(cur[i + 2]), 0.5f, b.z), h3 = fmaf(__uint_as_float(cur[i + 3]), 0.5f, b.w);
                const float hs0 = __sinf(h0), hc0 = __cosf(h0), hs1 = __sinf(h1), hc1 = __cosf(h1);
                const float hs2 = __sinf(h2), hc2 = __cosf(h2), hs3 = __sinf(h3), hc3 = __cosf(h3);
                const __half2 S01 = __floats2half2_rn(hs0, hs1), C01 = __floats2half2_rn(hc0, hc1);
                const __half2 S23 = __floats2half2_rn(hs2, hs3), C23 = __floats2half2_rn(hc2, hc3);
                const uint32_t q01 = cosq_enc2(S01, C01), q23 = cosq_enc2(S23, C23);
                cq[i / 4] = __byte_perm(q01, q23, 0x6420);              // codes sit in bytes 0 and 2 of each pair word
                if (last) {   // the output layer reads sin(pre) in fp32
                  s0 = (hs0 + hs0) * hc0; s1 = (hs1 + hs1) * hc1; s2 = (hs2 + hs2) * hc2; s3 = (hs3 + hs3) * hc3;
                  pk[i / 2] = pack_f16x2(s0, s1); pk[i / 2 + 1] = pack_f16x2(s2, s3);
                } else {
                  s0 = s1 = s2 = s3 = 0.f;
                  const __half2 p01 = __hmul2(S01, C01), p23 = __hmul2(S23, C23);
                  const __half2 H01 = __hadd2(p01, p01), H23 = __hadd2(p23, p23);
                  pk[i / 2] = *reinterpret_cast<const uint32_t *>(&H01); pk[i / 2 + 1] = *reinterpret_cast<const uint32_t *>(&H23);
                }
              } else {
                s0 = __sinf(__uint_as_float(cur[i]) + b.x); s1 = __sinf(__uint_as_float(cur[i + 1]) + b.y);
                s2 = __sinf(__uint_as_float(cur[i + 2]) + b.z); s3 = __sinf(__uint_as_float(cur[i + 3]) + b.w);
                pk[i / 2] = pack_f16x2(s0, s1); pk[i / 2 + 1] = pack_f16x2(s2, s3);
              }
              if (last) {   // fused output layer: out = W_out h + b_out
                const float4 wa = *reinterpret_cast<const float4 *>(wout_s + col0 + i);
                const float4 wb = *reinterpret_cast<const float4 *>(wout_s + D + col0 + i);
                o0 += s0 * wa.x + s1 * wa.y + s2 * wa.z + s3 * wa.w;
                o1 += s0 * wb.x + s1 * wb.y + s2 * wb.z + s3 * wb.w;
              }
            }
            if (TRAIN) {   // cos(pre) for the dgrad chain: CPT bytes per thread, a warp store covers 512 contiguous bytes
#pragma unroll
              for (int k = 0; k < CPT / 16; ++k)
                st_stream16(psave + (((sl * 4 + g * (CPT / 16) + k) * TILE_M + row) << 4), cq[4 * k], cq[4 * k + 1], cq[4 * k + 2], cq[4 * k + 3]);
            }
            if (!last || TRAIN) {
              if (h == 0) {
#pragma unroll
                for (int k = 0; k < CPT / 2; ++k) held[(CPT / 2) * j + k] = pk[k];
              } else {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                  *reinterpret_cast<uint4 *>(gA + sl * SLAB_BYTES + sw128_chunk_off(row, CHUNKS * g + c)) =
                      make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                fence_proxy_async_smem();
                if (!last) {
                  tcgen05_fence_before();
                  arrive_ready(1 + j);
                }
                if (TRAIN) arrive_wrote(sl);              // h_l for the weight gradients
              }
            }
          }
          if (h == 0 && (!last || TRAIN)) {
            // Half 0 of h_l goes into slabs 0..3 of the A image as soon as no MMA of layer l reads them any more - that is
            // half-way through the accumulation of half 1 (acc1a), normally already past: the 64 KB of shared-memory
            // stores and the hand-over to the issuer run under the remaining MMAs, and the next layer's first 16
            // instructions queue up right behind this layer's last one.
            mbar_wait(bar.acc1a(), ph);
            if (TRAIN) wait_afree();                  // ... and once the stores of h_{l-1} (or of the encoder image) have read it
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c = 0; c < CHUNKS; ++c)
                *reinterpret_cast<uint4 *>(gA + j * SLAB_BYTES + sw128_chunk_off(row, CHUNKS * g + c)) =
                    make_uint4(held[(CPT / 2) * j + 4 * c], held[(CPT / 2) * j + 4 * c + 1], held[(CPT / 2) * j + 4 * c + 2],
                               held[(CPT / 2) * j + 4 * c + 3]);
            fence_proxy_async_smem();
            if (!last) {   // hand the slabs to the MMA issuer first: the stores are off the critical path
              tcgen05_fence_before();
              arrive_ready(0);
            }
            if (TRAIN) arrive_wrote(0);
          }
        }
        ph ^= 1;
        if (last) {
          tcgen05_fence_before();
          arrive_ready(4);                            // D half 1 drained: the next tile's layer 0 may use it
          // combine the column groups of each row
          if (g > 0) osum_s[(g - 1) * TILE_M + row] = make_float2(o0, o1);
          named_bar_sync(1, N_EPI);
          if (g == 0 && m < p.M) {
#pragma unroll
            for (int gg = 1; gg < EPI_GROUPS; ++gg) { const float2 o = osum_s[(gg - 1) * TILE_M + row]; o0 += o.x; o1 += o.y; }
            p.out[m] = make_float2(o0 + __ldg(b_out) + p.off0, o1 + __ldg(b_out + 1) + p.off1);
          }
        }
      }
    }
#ifdef SNF_PROF
    if (e == 0 && lane == 0) {
      g_prof[TRAIN][blockIdx.x * 8 + 3] = clock64() - t_begin;
      g_prof[TRAIN][blockIdx.x * 8 + 4] = t_acc0;
      g_prof[TRAIN][blockIdx.x * 8 + 5] = t_acc1;
      g_prof[TRAIN][blockIdx.x * 8 + 6] = t_enc;
      g_prof[TRAIN][blockIdx.x * 8 + 7] = t_stw;
    }
#endif
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
                      float *const *gB, int num_sms, int wsplit, cudaStream_t st);

int64_t snf_mlp_bf16_ws_bytes(int64_t M, int train, int x3) { return bf::bf16_layout(nullptr, M, train, x3).bytes; }

#ifdef SNF_PROF
// debug build only: per-CTA cycle counters of the last forward launches ([0] inference, [1] training)
extern "C" int snf_debug_prof(unsigned long long *host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, bf::g_prof, sizeof(bf::g_prof));
}
extern "C" int snf_debug_trace(long long *host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, bf::g_trace, sizeof(bf::g_trace));
}
#endif

extern "C" int64_t snf_mlp_pack_bytes(void) { return bf::PACK_TOTAL_BYTES; }

extern "C" int snf_mlp_pack_bf16(const float *const *W, const float *const *B, void *packed, void *stream) {
  SNF_CHECK_PTR(W); SNF_CHECK_PTR(B); SNF_CHECK_PTR(packed); SNF_CHECK_ALIGN(packed, 1024);
  for (int l = 0; l <= bf::NH; ++l) { SNF_CHECK_PTR(W[l]); SNF_CHECK_PTR(B[l]); }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = (int64_t)bf::FWD_BLOCKS * (bf::WBLK_BYTES / 16);
  bf::pack_weights_kernel<<<(unsigned)ceil_div64(chunks, 256), 256, 0, st>>>(
      W[0], W[1], W[2], W[3], W[4], W[5], W[6], W[7], reinterpret_cast<uint4 *>(packed),
      reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_LO_OFF));
  const int nsmall = bf::NH * bf::D + 2 * bf::D + 2;
  bf::pack_small_kernel<<<(nsmall + 255) / 256, 256, 0, st>>>(B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], W[8], B[8],
                                                             reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_BIAS_OFF));
  count_launch(2);
  return snf_bf16_pack_wt(W, packed, st);   // W^T blocks for the dgrad chain
}

// Per-DEVICE one-time setup (the reference renders through nn.DataParallel from a thread pool, evaluation/loader.py:
// 37-39,143,226-229: several devices and several host threads in one process).  cudaFuncAttributeMaxDynamicSharedMemorySize
// and the SM count belong to the current device, not to the process.
int snf_device_setup(int *num_sms_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= snf::kMaxDevices) return SNF_E_ARG;
  snf::DeviceState &d = snf::g_devices[dev];
  std::call_once(d.once, [&]() {
    int n = 0;
    cudaError_t err = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    d.num_sms = (err == cudaSuccess && n > 0) ? n : 148;
    d.status = snf_set_kernel_attributes();
  });
  if (num_sms_out) {
    const int r = __atomic_load_n(&d.reserve, __ATOMIC_RELAXED);
    int n = d.num_sms - r;
    *num_sms_out = n >= 2 ? n : 2;
  }
  return d.status;
}

// Multi-GPU training: the persistent field-network kernels fill every SM with one 224 KB CTA, so a concurrent NCCL kernel
// finds no SM until a grid drains - and once its CTAs sit on a few SMs, spinning for the peer, the next persistent grid
// cannot place its CTAs there.  Leaving a few SMs out of the persistent grids gives the collective a home of its own.
extern "C" int snf_config_reserve_sms(int n) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= snf::kMaxDevices || n < 0 || n > 64) return SNF_E_ARG;
  __atomic_store_n(&snf::g_devices[dev].reserve, n & ~1, __ATOMIC_RELAXED);   // whole CTA pairs
  return 0;
}

int snf_bf16_set_attributes_bwd();   // snf_mlp_bf16_bwd.cu
int snf_sampling_set_attributes();   // snf_sampling.cu
int snf_x3_set_attributes();         // snf_mlp_x3.cu
int snf_set_kernel_attributes() {
  cudaError_t e = cudaFuncSetAttribute(bf::mlp_fwd_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::fw::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(bf::mlp_fwd_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::fw::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  if (int r = snf_bf16_set_attributes_bwd()) return r;
  if (int r = snf_x3_set_attributes()) return r;
  return snf_sampling_set_attributes();
}

extern "C" int snf_mlp_fwd_bf16(const float *x, int64_t M, const void *packed, float off0, float off1, float *out,
                                void *ws, int train, void *stream) {
  if (M == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(x); SNF_CHECK_PTR(packed); SNF_CHECK_PTR(out);
  SNF_CHECK_ALIGN(x, 16); SNF_CHECK_ALIGN(out, 8); SNF_CHECK_ALIGN(packed, 1024);
  if (M < 0) return SNF_E_ARG;
  if (train) { SNF_CHECK_PTR(ws); SNF_CHECK_ALIGN(ws, 1024); }
  int nsm = 0;
  if (int e = snf_device_setup(&nsm)) return e;
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
  int grid = p.num_tiles < nsm ? p.num_tiles : nsm;
  grid &= ~1;   // whole CTA pairs
  if (train) bf::mlp_fwd_bf16_kernel<true><<<grid, bf::NTHREADS, bf::fw::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  else bf::mlp_fwd_bf16_kernel<false><<<grid, bf::NTHREADS, bf::fw::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  count_launch();
  return launch_status();
}

static int mlp_bwd_16(int64_t M, const void *packed, const float *grad_out, void *ws, float *const *gW, float *const *gB,
                      int wsplit, void *stream);

extern "C" int snf_mlp_bwd_bf16(const float *x, int64_t M, const void *packed, const float *grad_out, void *ws,
                                float *const *gW, float *const *gB, void *stream) {
  (void)x;
  return mlp_bwd_16(M, packed, grad_out, ws, gW, gB, 1, stream);
}

// backward of snf_mlp_fwd_x3: the 16-bit kernels on the saved high halves, the dgrad chain's W^T operand as (hi, lo)
extern "C" int snf_mlp_bwd_x3(const float *x, int64_t M, const void *packed, const float *grad_out, void *ws,
                              float *const *gW, float *const *gB, void *stream) {
  (void)x;
  return mlp_bwd_16(M, packed, grad_out, ws, gW, gB, 2, stream);
}

static int mlp_bwd_16(int64_t M, const void *packed, const float *grad_out, void *ws, float *const *gW, float *const *gB,
                      int wsplit, void *stream) {
  SNF_CHECK_PTR(packed); SNF_CHECK_PTR(grad_out); SNF_CHECK_PTR(ws); SNF_CHECK_PTR(gW); SNF_CHECK_PTR(gB);
  SNF_CHECK_ALIGN(grad_out, 8); SNF_CHECK_ALIGN(ws, 1024); SNF_CHECK_ALIGN(packed, 1024);
  if (M <= 0) return SNF_E_ARG;
  for (int l = 0; l <= bf::NH; ++l) { SNF_CHECK_PTR(gW[l]); SNF_CHECK_PTR(gB[l]); SNF_CHECK_ALIGN(gW[l], 16); }
  bf::Bf16Ws w = bf::bf16_layout(ws, M, 1);
  int nsm = 0;
  if (int e = snf_device_setup(&nsm)) return e;
  return snf_bf16_backward(grad_out, M, packed, w, gW, gB, nsm, wsplit, (cudaStream_t)stream);
}
