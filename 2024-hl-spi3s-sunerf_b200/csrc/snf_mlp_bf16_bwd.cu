// 16-bit tensor-core backward of the field network (sm_100a).  Two persistent warp-specialised kernels over the
// tile images the forward saved (snf_mlp_bf16.cu: enc, H_l = sin(pre_l) as [tile][..][128 x 512] fp16 in the UMMA K-major
// SWIZZLE_128B image, C_l = cos(pre_l) as one-byte codes).  All gradients images carry the power-of-two scale S of
// GradScale (snf_bf16_common.cuh): fp16 operands need it for range, the wgrad flush divides it out exactly.
//
//  dgrad chain  (mlp_dgrad_bf16_kernel): per 128-point tile, dpre_7 = S (g W_out) * C_7 in the epilogue warps, then for
//     l = 7..1:  dpre_{l-1} = (dpre_l W_l) * C_{l-1}  -- tcgen05.mma with A = dpre_l image in shared memory, B = W_l^T
//     blocks streamed by TMA, D in TMEM; every dpre_l image is bulk-stored to HBM (D_l) for the weight gradients.
//  wgrad        (mlp_wgrad_bf16_kernel): dW_l[o,i] = sum_p D_l[p,o] Hprev_l[p,i].  The saved images are read
//     "transposed" as MN-major UMMA operands (same bytes, different descriptor), 64 points per pipeline stage;
//     a CTA owns a 128(o) x 512(i) fp32 accumulator (all of TMEM) for a range of tiles, then flushes it with
//     red.global.add.v4.f32 into the flat gradient buffer.  The otherwise idle epilogue warps sum the D_l stages
//     over points for the bias gradients.
//
// This is what torch.autograd derives for NeRF.forward (sunerf/model/model.py:44-57) - restated analytically.
#include <cstdio>
#include <cstdlib>
#include "snf_bf16_common.cuh"

namespace snf {
namespace bf {

constexpr int WG_THREADS = 192;   // wgrad: producer warp, MMA warp, 4 bias/flush warps

// ------------------------------------------------------------------------------------------ W^T packing
// block (l, nh, ks): rows = input feature i (256 per block), k = output feature o (64 per block): B[i][o] = W_l[o][i]
__global__ void __launch_bounds__(256) pack_wt_kernel(const float *w1, const float *w2, const float *w3, const float *w4,
                                                      const float *w5, const float *w6, const float *w7,
                                                      uint4 *__restrict__ dst, uint4 *__restrict__ dst_lo) {
  const float *W[7] = {w1, w2, w3, w4, w5, w6, w7};
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)WT_BLOCKS * (WBLK_BYTES / 16)) return;
  const int blk = (int)(idx / (WBLK_BYTES / 16)), within = (int)(idx % (WBLK_BYTES / 16));
  const int li = blk / 16, q = (blk % 16) >> 3, ks = blk & 7;
  const int r = within >> 3, pos = within & 7, c8 = pos ^ (r & 7);
  const int i = q * NCHUNK + r, o0 = ks * 64 + c8 * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = W[li][(o0 + j) * D + i];
  uint4 o;
  o.x = pack_f16x2(v[0], v[1]); o.y = pack_f16x2(v[2], v[3]); o.z = pack_f16x2(v[4], v[5]); o.w = pack_f16x2(v[6], v[7]);
  dst[(int64_t)blk * (WBLK_BYTES / 16) + within] = o;
  float lo[8];                                  // low halves: the split-precision dgrad chain (WSPLIT = 2)
#pragma unroll
  for (int j = 0; j < 8; ++j) lo[j] = v[j] - __half2float(__float2half_rn(v[j]));
  o.x = pack_f16x2(lo[0], lo[1]); o.y = pack_f16x2(lo[2], lo[3]); o.z = pack_f16x2(lo[4], lo[5]); o.w = pack_f16x2(lo[6], lo[7]);
  dst_lo[(int64_t)blk * (WBLK_BYTES / 16) + within] = o;
}

// ------------------------------------------------------------------------------------------ dgrad chain
struct DgradParams {
  const float2 *g;          // [M] dL/d out
  int64_t M;
  int num_tiles;            // even
  const uint8_t *packed;    // forward pack + W^T blocks
  const uint8_t *save_pre;  // [tiles][8][64 KB] cos(pre) codes, C_BYTES layout (written by the forward)
  uint8_t *save_d;          // [tiles][8][128 KB] S dpre_l images, fp16 (output)
  GradScale *scale;         // in: gmax_bits (absmax_kernel); out: S, 1/S (written by CTA 0 for the wgrad)
};

// max |g| over the whole batch as the bit pattern of a non-negative float (ordered like an unsigned; a NaN wins)
__global__ void __launch_bounds__(256) absmax_kernel(const float2 *__restrict__ g, int64_t M, uint32_t *__restrict__ out) {
  uint32_t m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = __ldg(g + i);
    m = max(m, max(__float_as_uint(fabsf(v.x)), __float_as_uint(fabsf(v.y))));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(kFull, m, d));
  if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(out, m);
}

__device__ __forceinline__ void prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// dpre_{l-1} = (dpre_l W_l) * cos(pre_{l-1}); the cosines come from the forward as one-byte codes (C_BYTES images).
// Same overlapped structure as the forward (snf_mlp_bf16.cu): per layer two temporal N-halves, the epilogue of half 0
// runs under the MMAs of half 1 and keeps its result in registers until the A image may be overwritten; the next
// layer's MMAs start slab by slab.  Shared-memory layout and barriers: namespace fw (the bias area is unused).
// WSPLIT = 2 (backward of the split-precision mode): W^T enters as the pair (hi, lo) - the same layer is accumulated over
// sixteen weight blocks instead of eight, the second eight against the low halves and the same A slabs.
template <int WSPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) mlp_dgrad_bf16_kernel(const DgradParams p) {
  constexpr int NSTAGE = fw::NSTAGE, OFF_RING = fw::OFF_RING, OFF_WOUT = fw::OFF_WOUT, OFF_BAR = fw::OFF_BAR;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  uint8_t *gA = smem_raw;
  const uint32_t sA = base, sW = base + OFF_RING;
  float *wout_s = reinterpret_cast<float *>(smem_raw + OFF_WOUT);          // [2][512]
  const fw::Bars bar{base + OFF_BAR};
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar.full(s), rank == 0 ? 2 : 1); mbar_init(bar.empty(s), 1); }
    mbar_init(bar.acc(0), 1); mbar_init(bar.acc(1), 1);
    for (int k = 0; k < 5; ++k) mbar_init(bar.ready(k), 2 * N_EPI_WARPS);
    for (int k = 0; k < 8; ++k) mbar_init(bar.wrote(k), N_EPI_WARPS);
    mbar_init(bar.afree(), 1);
    mbar_init(bar.acc1a(), 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * D; i += NTHREADS) wout_s[i] = __ldg(reinterpret_cast<const float *>(p.packed + PACK_WOUT_OFF) + i);
  if (warp == 1) tmem_alloc_2cta(bar.tmem_slot(), 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem_raw + fw::TMEM_SLOT_OFF);
  const uint8_t *wt = p.packed + PACK_WT_OFF, *wt_lo = p.packed + PACK_WTLO_OFF;
  constexpr int NB = 16 * WSPLIT;               // weight blocks per layer: (part, N-half, k-slab), the N-half outermost

  if (warp < EPI_WARP0) {
  reg_dealloc<REGS_CTRL>();
  if (warp == 0) {
    // =========================== TMA producer: W^T blocks through the ring, plus L2 prefetches of the saved
    // pre-activation images about one layer ahead of the epilogue that multiplies by their cosine
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const uint64_t keep = l2_policy_evict_last();
      auto pre_img = [&](int tile, int l) { return p.save_pre + ((int64_t)tile * NH + l) * C_BYTES; };
      {
        const int tile = pair * 2 + (int)rank;
        if (tile < p.num_tiles)
          for (int sl = 0; sl < 8; ++sl) {
            prefetch_l2(pre_img(tile, NH - 1) + sl * (C_BYTES / 8), C_BYTES / 8);
            prefetch_l2(pre_img(tile, NH - 2) + sl * (C_BYTES / 8), C_BYTES / 8);
          }
      }
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        const int tile = tp * 2 + (int)rank, next_tile = tile + 2 * npairs;
        for (int l = NH - 1; l >= 1; --l)
          for (int bb = 0; bb < NB; ++bb) {
            // consumption order: N-half h, then part (hi, lo), then k-slab
            const int h = bb / (8 * WSPLIT), part = (bb / 8) % WSPLIT, ks = bb & 7;
            const int b = h * 8 + ks;                                        // block index inside the layer's 16 blocks
            mbar_wait(bar.empty(s), ph ^ 1);
            mbar_arrive_expect_tx(bar.full(s), WHALF_BYTES);
            bulk_g2s_hint(sW + s * WHALF_BYTES, (part ? wt_lo : wt) + (int64_t)((l - 1) * 16 + b) * WBLK_BYTES + rank * WHALF_BYTES,
                          WHALF_BYTES, bar.full(s), keep);
            if (part == 0) {
              if (l >= 2) {
                if ((b & 1) == 0) prefetch_l2(pre_img(tile, l - 2) + (b >> 1) * (C_BYTES / 8), C_BYTES / 8);
              } else if (next_tile < p.num_tiles) {   // l == 1: the next tile's first two images
                prefetch_l2(pre_img(next_tile, b < 8 ? NH - 1 : NH - 2) + (b & 7) * (C_BYTES / 8), C_BYTES / 8);
              }
            }
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    if (rank == 0) {
      // =========================== MMA issuer (leader CTA); see the forward kernel
      const uint32_t idesc = idesc_f16kind(256, NCHUNK, FMT, FMT);
      const uint64_t adesc0 = smem_desc(sA, 16, 1024), bdesc0 = smem_desc(sW, 16, 1024);
      uint32_t rph = 0;
      auto wait_ready = [&](int k) { mbar_wait(bar.ready(k), (rph >> k) & 1u); rph ^= 1u << k; };
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        for (int l = NH - 1; l >= 1; --l) {
          for (int h = 0; h < 2; ++h) {
            for (int kk = 0; kk < 8 * WSPLIT; ++kk) {
              const int ks = kk & 7;                    // A slab; kk >= 8: the same slabs against the low halves of W^T
              if (h == 0 && kk < 8) {
                if (ks == 0) wait_ready(0);
                else if (ks >= 4) wait_ready(ks - 3);
              }
              mbar_wait(bar.full(s), ph);
              tcgen05_fence_after();
              if (elect_one()) {
                const uint64_t ad = adesc0 + (uint64_t)((ks * SLAB_BYTES) >> 4), bd = bdesc0 + (uint64_t)((s * WHALF_BYTES) >> 4);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) mma_ss_2cta(tmem + h * NCHUNK, ad + 2 * k4, bd + 2 * k4, idesc, (kk | k4) != 0);
                mma_commit_2cta(bar.empty(s), 3);
                if (kk == 8 * WSPLIT - 1) mma_commit_2cta(bar.acc(h), 3);
                if (h == 1 && kk == 8 * (WSPLIT - 1) + 3) mma_commit_2cta(bar.acc1a(), 3);   // slabs 0..3 of A are no longer read
              }
              __syncwarp();
              if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    } else if (lane == 0) {
      for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
        for (int blk = 0; blk < WT_BLOCKS * WSPLIT; ++blk) {
          mbar_wait(bar.full(s), ph);
          mbar_arrive_remote_relaxed(mapa_shared(bar.full(s), 0));   // the data is TMA-written and tensor-core-read
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2 && lane == 0) {
    // =========================== store warp: TMA-stores every finished slab of the A image (dpre_7 .. dpre_0) for the
    // wgrad; the epilogue warps only arrive on wrote[] (see the forward kernel)
    const uint64_t stream_pol = l2_policy_evict_first();
    uint32_t wph = 0;
    auto wait_wrote = [&](int slot) { mbar_wait(bar.wrote(slot), (wph >> slot) & 1u); wph ^= 1u << slot; };
    auto store_slabs = [&](uint8_t *img, int sl0, int nsl) {
      for (int sl = sl0; sl < sl0 + nsl; ++sl) bulk_s2g_hint(img + sl * SLAB_BYTES, sA + sl * SLAB_BYTES, SLAB_BYTES, stream_pol);
      bulk_commit();
    };
    for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
      const int tile = tp * 2 + (int)rank;
      uint8_t *d_tile = p.save_d + (int64_t)tile * NH * A_BYTES;
      for (int sl = 0; sl < 8; ++sl) { wait_wrote(sl); store_slabs(d_tile + (int64_t)(NH - 1) * A_BYTES, sl, 1); }
      bulk_wait_read_all();
      mbar_arrive(bar.afree());                         // dpre_7 has left the A image
      for (int l = NH - 1; l >= 1; --l) {
        uint8_t *dprev = d_tile + (int64_t)(l - 1) * A_BYTES;
        wait_wrote(0);
        store_slabs(dprev, 0, 4);
        for (int j = 0; j < 4; ++j) { wait_wrote(4 + j); store_slabs(dprev, 4 + j, 1); }
        bulk_wait_read_all();
        mbar_arrive(bar.afree());
      }
    }
    bulk_wait_all();
  }
  } else {
    reg_alloc<REGS_EPI>();
    // =========================== epilogue warps: thread = (row, g); step (h, j) <-> slab 4h + j, chunks CHUNKS g ..
    const int e = warp - EPI_WARP0, q = warp & 3, g = e >> 2, row = q * 32 + lane;
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16) + g * CPT;
    uint32_t ready_addr[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ready_addr[k] = rank == 0 ? bar.ready(k) : mapa_shared(bar.ready(k), 0);
    auto arrive_ready = [&](int k) {
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(ready_addr[k]);
        else mbar_arrive_remote_relaxed(ready_addr[k]);
      }
    };
    // the scale of this backward call: the same S in every CTA (same inputs, same arithmetic); CTA 0 publishes it
    float gscale;
    {
      float wm = 0.f;
      for (int j = lane; j < D; j += 32) wm = fmaxf(wm, fabsf(wout_s[j]) + fabsf(wout_s[D + j]));
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) wm = fmaxf(wm, __shfl_xor_sync(kFull, wm, d));
      gscale = grad_scale_from(__uint_as_float(__ldg(&p.scale->gmax_bits)), wm);
      if (blockIdx.x == 0 && e == 0 && lane == 0) { p.scale->S = gscale; p.scale->invS = 1.f / gscale; }
    }
    // this thread's CPT cosine codes of one step: CPT / 16 x 16 bytes (cosq_dec)
    constexpr int NQ = CPT / 16;
    auto load_pre = [&](const uint8_t *img, int sl, uint4 (&pv)[NQ]) {
#pragma unroll
      for (int k = 0; k < NQ; ++k)
        pv[k] = __ldcs(reinterpret_cast<const uint4 *>(img + (((sl * 4 + g * NQ + k) * TILE_M + row) << 4)));
    };
    // (a0 cos_0, a1 cos_1) as an fp16 pair for the codes 2 pr, 2 pr + 1 of the word w: the products with |cos| are formed in
    // fp32 and rounded once, the two signs are flipped on the packed pair
#define SNF_MUL2(a0, a1, w, pr) \
    (pack_f16x2((a0) * cosq_dec_abs<2 * (pr)>(w), (a1) * cosq_dec_abs<2 * (pr) + 1>(w)) ^ cosq_sign2<pr>(w))
    // hand complete slabs to the MMA issuer first (ready barrier k, or none), then to the store warp (wrote[slot])
    auto publish = [&](int slot, int k) {
      fence_proxy_async_smem();
      if (k >= 0) { tcgen05_fence_before(); arrive_ready(k); }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar.wrote(slot));
    };
    uint32_t aph = 0;
    bool first_tile = true;
    auto wait_afree = [&]() { mbar_wait(bar.afree(), aph); aph ^= 1; };   // the previous stores have read the A image
    uint32_t ph = 0;
    for (int tp = pair; tp * 2 < p.num_tiles; tp += npairs) {
      const int tile = tp * 2 + (int)rank;
      const int64_t m = (int64_t)tile * TILE_M + row;
      const uint8_t *pre_tile = p.save_pre + (int64_t)tile * NH * C_BYTES;
      uint4 pn[NQ];
      // ---- dpre_7 = (g0 W_out[0,:] + g1 W_out[1,:]) * cos(pre_7), written straight into the A image
      {
        const uint8_t *p7 = pre_tile + (int64_t)(NH - 1) * C_BYTES;
        load_pre(p7, 0, pn);
        float2 gg = make_float2(0.f, 0.f);
        if (m < p.M) { gg = p.g[m]; gg.x *= gscale; gg.y *= gscale; }
        if (!first_tile) wait_afree();                // the previous tile's last stores have left the A image
        first_tile = false;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
          uint4 pv[NQ];
#pragma unroll
          for (int k = 0; k < NQ; ++k) pv[k] = pn[k];
          if (sl + 1 < 8) load_pre(p7, sl + 1, pn);
          else load_pre(pre_tile + (int64_t)(NH - 2) * C_BYTES, 0, pn);
#pragma unroll
          for (int c = 0; c < CHUNKS; ++c) {
            const int col = sl * 64 + (CHUNKS * g + c) * 8;
            const float4 wa0 = *reinterpret_cast<const float4 *>(wout_s + col), wa1 = *reinterpret_cast<const float4 *>(wout_s + col + 4);
            const float4 wb0 = *reinterpret_cast<const float4 *>(wout_s + D + col), wb1 = *reinterpret_cast<const float4 *>(wout_s + D + col + 4);
            const uint4 &cv = pv[c >> 1];             // 8 codes of this chunk: two words
            const uint32_t cw0 = (c & 1) ? cv.z : cv.x, cw1 = (c & 1) ? cv.w : cv.y;
            uint4 o;
            o.x = SNF_MUL2(gg.x * wa0.x + gg.y * wb0.x, gg.x * wa0.y + gg.y * wb0.y, cw0, 0);
            o.y = SNF_MUL2(gg.x * wa0.z + gg.y * wb0.z, gg.x * wa0.w + gg.y * wb0.w, cw0, 1);
            o.z = SNF_MUL2(gg.x * wa1.x + gg.y * wb1.x, gg.x * wa1.y + gg.y * wb1.y, cw1, 0);
            o.w = SNF_MUL2(gg.x * wa1.z + gg.y * wb1.z, gg.x * wa1.w + gg.y * wb1.w, cw1, 1);
            *reinterpret_cast<uint4 *>(gA + sl * SLAB_BYTES + sw128_chunk_off(row, CHUNKS * g + c)) = o;
          }
          publish(sl, sl == 3 ? 0 : sl >= 4 ? sl - 3 : -1);
        }
      }
#pragma unroll 1
      for (int l = NH - 1; l >= 1; --l) {
        // accumulator = dpre_l W_l = dL/dh_{l-1}; multiply by cos(pre_{l-1}) -> dpre_{l-1}
        const bool last = (l == 1);
        const uint8_t *pprev = pre_tile + (int64_t)(l - 1) * C_BYTES;
        const uint8_t *pnext = pre_tile + (int64_t)(l >= 2 ? l - 2 : 0) * C_BYTES;
        uint32_t held[4 * CPT / 2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar.acc(h), ph);
          tcgen05_fence_after();
          uint32_t accA[16], accB[16];                // 16-column TMEM loads, double buffered
          tmem_ld16(tm_row + h * 256, accA);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int sl = h * 4 + j;
            uint4 pv[NQ];
#pragma unroll
            for (int k = 0; k < NQ; ++k) pv[k] = pn[k];
            if (sl + 1 < 8) load_pre(pprev, sl + 1, pn);
            else if (!last) load_pre(pnext, 0, pn);
            uint32_t pk[CPT / 2];
            auto part16 = [&](const uint32_t (&a)[16], int c0) {   // 16 columns = chunks c0, c0 + 1 of this step
              const uint4 &cv = pv[c0 >> 1];
              const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const uint32_t w4 = cw[i >> 2];
                pk[4 * c0 + i / 2] = SNF_MUL2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]), w4, 0);
                pk[4 * c0 + i / 2 + 1] = SNF_MUL2(__uint_as_float(a[i + 2]), __uint_as_float(a[i + 3]), w4, 1);
              }
            };
            if (CPT == 32) {
              tmem_ld_wait(accA);
              tmem_ld16(tm_row + h * 256 + j * 64 + 16, accB);
              part16(accA, 0);
              tmem_ld_wait(accB);
              if (j + 1 < 4) tmem_ld16(tm_row + h * 256 + (j + 1) * 64, accA);
              part16(accB, 2);
            } else {                                  // CPT == 16: one load per step, the next step's in flight
              uint32_t(&cur)[16] = (j & 1) ? accB : accA;
              uint32_t(&nxt)[16] = (j & 1) ? accA : accB;
              tmem_ld_wait(cur);
              if (j + 1 < 4) tmem_ld16(tm_row + h * 256 + (j + 1) * 64, nxt);
              part16(cur, 0);
            }
            if (h == 0) {
#pragma unroll
              for (int k = 0; k < CPT / 2; ++k) held[(CPT / 2) * j + k] = pk[k];
            } else {
#pragma unroll
              for (int c = 0; c < CHUNKS; ++c)
                *reinterpret_cast<uint4 *>(gA + sl * SLAB_BYTES + sw128_chunk_off(row, CHUNKS * g + c)) =
                    make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
              publish(sl, last ? -1 : 1 + j);
            }
          }
          if (h == 0) {
            // half 0 of dpre_{l-1} goes into slabs 0..3 of the A image as soon as no MMA of layer l reads them any more
            // (acc1a: half-way through the accumulation of half 1) - see the forward kernel
            mbar_wait(bar.acc1a(), ph);
            wait_afree();                             // ... and once the stores of dpre_l have read it
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c = 0; c < CHUNKS; ++c)
                *reinterpret_cast<uint4 *>(gA + j * SLAB_BYTES + sw128_chunk_off(row, CHUNKS * g + c)) =
                    make_uint4(held[(CPT / 2) * j + 4 * c], held[(CPT / 2) * j + 4 * c + 1], held[(CPT / 2) * j + 4 * c + 2],
                               held[(CPT / 2) * j + 4 * c + 3]);
            publish(0, last ? -1 : 0);
          }
        }
        ph ^= 1;
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

// ------------------------------------------------------------------------------------------ wgrad
// CTA pairs (tcgen05 cta_group::2, M = 256 output features across the pair, N = 256 input features per instruction):
// CTA r of a pair owns output features [256 ob + 128 r, +128) - its A operand (2 slabs of D_l) and its 128 x 512 fp32
// accumulator (all of its TMEM) - and supplies half of every B operand (2 of the 4 H slabs of each N = 256
// instruction), so a pipeline stage of 64 points is 48 KB per CTA instead of the 80 KB a single-CTA tile needs.
constexpr int WG_KSTAGE = 64;                           // points per pipeline stage
constexpr int WG_SLAB_STAGE = WG_KSTAGE * 128;          // 8 KB: 64 rows of one 64-feature slab
constexpr int WG_STAGE_BYTES = 6 * WG_SLAB_STAGE;       // 2 (own o) + 2 x 2 (own half of i, two N-halves) slabs = 48 KB
constexpr int WG_NSTAGE = 4;   // few large copies: a single producer thread sustains only ~1 bulk copy per 100-300 cycles
constexpr int WG_SMEM_BYTES = WG_NSTAGE * WG_STAGE_BYTES + 256;
static_assert(WG_SMEM_BYTES <= 232448, "wgrad ring exceeds shared memory");

struct WgradParams {
  const uint8_t *save_d, *save_h, *save_enc;
  const GradScale *scale;   // 1/S written by the dgrad chain
  int num_tiles, tiles_per_item, num_items;
  float *gW[NH];
  float *gB[NH];
};

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1) mlp_wgrad_bf16_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  uint8_t *gS = smem_raw;
  const uint32_t sBar = base + WG_NSTAGE * WG_STAGE_BYTES;
  auto bar_full = [&](int s) { return sBar + 8u * s; };                       // leader: own TMA + peer relay
  auto bar_empty = [&](int s) { return sBar + 8u * (WG_NSTAGE + s); };       // multicast commit + this CTA's 4 bias warps
  const uint32_t bar_acc = sBar + 8u * (2 * WG_NSTAGE);                      // item accumulated (multicast commit)
  const uint32_t bar_accfree = sBar + 8u * (2 * WG_NSTAGE + 1);              // leader: 8 flush warps (both CTAs) done
  const uint32_t tmem_slot = sBar + 8u * (2 * WG_NSTAGE + 2);
  volatile uint32_t *tmem_slot_g = reinterpret_cast<volatile uint32_t *>(gS + WG_NSTAGE * WG_STAGE_BYTES + 8 * (2 * WG_NSTAGE + 2));
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_NSTAGE; ++s) { mbar_init(bar_full(s), rank == 0 ? 2 : 1); mbar_init(bar_empty(s), 1 + 4); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_accfree, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot_g;

  // item -> (tile range, layer, 256-wide o-block)
  auto decode = [&](int item, int &l, int &ob, int &t0, int &t1) {
    ob = item & 1; l = (item >> 1) & 7;
    const int r = item >> 4;
    t0 = r * p.tiles_per_item;
    t1 = min(t0 + p.tiles_per_item, p.num_tiles);
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int item = pair; item < p.num_items; item += npairs) {
        int l, ob, t0, t1; decode(item, l, ob, t0, t1);
        const int nb = l == 0 ? 1 : 4;
        for (int t = t0; t < t1; ++t) {
          const uint8_t *dimg = p.save_d + ((int64_t)t * NH + l) * A_BYTES + (int64_t)(ob * 4 + rank * 2) * SLAB_BYTES;
          const uint8_t *himg = l == 0 ? p.save_enc + (int64_t)t * 2 * SLAB_BYTES
                                       : p.save_h + ((int64_t)t * NH + (l - 1)) * A_BYTES;
          for (int qd = 0; qd < TILE_M / WG_KSTAGE; ++qd) {
            mbar_wait(bar_empty(s), ph ^ 1);
            mbar_arrive_expect_tx(bar_full(s), (2 + nb) * WG_SLAB_STAGE);
            const uint32_t st = base + s * WG_STAGE_BYTES;
            for (int j = 0; j < 2; ++j)
              bulk_g2s(st + j * WG_SLAB_STAGE, dimg + j * SLAB_BYTES + qd * WG_SLAB_STAGE, WG_SLAB_STAGE, bar_full(s));
            if (l == 0) {
              bulk_g2s(st + 2 * WG_SLAB_STAGE, himg + rank * SLAB_BYTES + qd * WG_SLAB_STAGE, WG_SLAB_STAGE, bar_full(s));
            } else {
              for (int nh = 0; nh < 2; ++nh)
                for (int j = 0; j < 2; ++j)
                  bulk_g2s(st + (2 + nh * 2 + j) * WG_SLAB_STAGE, himg + (nh * 4 + rank * 2 + j) * SLAB_BYTES + qd * WG_SLAB_STAGE,
                           WG_SLAB_STAGE, bar_full(s));
            }
            if (++s == WG_NSTAGE) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    if (rank == 0) {
      // MMA issuer: the whole warp runs the uniform control flow, one elected lane issues
      const uint32_t idesc256 = idesc_f16kind(256, 256, FMT, FMT, 1, 1), idesc128 = idesc_f16kind(256, 128, FMT, FMT, 1, 1);
      const uint64_t desc0 = smem_desc(base, WG_SLAB_STAGE, 1024);   // MN-major: LBO = slab stride in the stage, SBO = 8-point groups
      uint32_t ph_free = 0;
      bool first_item = true;
      for (int item = pair; item < p.num_items; item += npairs) {
        int l, ob, t0, t1; decode(item, l, ob, t0, t1);
        if (!first_item) { mbar_wait(bar_accfree, ph_free); ph_free ^= 1; tcgen05_fence_after(); }
        first_item = false;
        uint32_t accumulate = 0;
        for (int t = t0; t < t1; ++t)
          for (int qd = 0; qd < TILE_M / WG_KSTAGE; ++qd) {
            mbar_wait(bar_full(s), ph);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint64_t sd = desc0 + (uint64_t)((s * WG_STAGE_BYTES) >> 4);
#pragma unroll
              for (int k16 = 0; k16 < WG_KSTAGE / 16; ++k16) {
                const uint64_t ad = sd + (uint64_t)((k16 * 2048) >> 4);
                if (l == 0) {
                  mma_ss_2cta(tmem, ad, ad + (uint64_t)((2 * WG_SLAB_STAGE) >> 4), idesc128, accumulate | k16);
                } else {
#pragma unroll
                  for (int nh = 0; nh < 2; ++nh)
                    mma_ss_2cta(tmem + nh * 256, ad, ad + (uint64_t)(((2 + nh * 2) * WG_SLAB_STAGE) >> 4), idesc256, accumulate | k16);
                }
              }
              mma_commit_2cta(bar_empty(s), 3);
              if (t == t1 - 1 && qd == TILE_M / WG_KSTAGE - 1) mma_commit_2cta(bar_acc, 3);
            }
            __syncwarp();
            accumulate = 1;
            if (++s == WG_NSTAGE) { s = 0; ph ^= 1; }
          }
      }
    } else if (lane == 0) {
      // peer relay: this CTA's part of the stage has landed
      for (int item = pair; item < p.num_items; item += npairs) {
        int l, ob, t0, t1; decode(item, l, ob, t0, t1);
        for (int i = (t1 - t0) * (TILE_M / WG_KSTAGE); i > 0; --i) {
          mbar_wait(bar_full(s), ph);
          mbar_arrive_remote_relaxed(mapa_shared(bar_full(s), 0));
          if (++s == WG_NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // bias-gradient partial sums while the pipeline runs, accumulator flush at the end of each item
    const int q = warp & 3;                             // thread <-> output feature o = 256 ob + 128 rank + row
    const int row = q * 32 + lane;
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t accfree_addr = rank == 0 ? bar_accfree : mapa_shared(bar_accfree, 0);
    int s = 0; uint32_t ph = 0, ph_acc = 0;
    const float inv_s = __ldg(&p.scale->invS);          // the D images hold S dL/dpre
    for (int item = pair; item < p.num_items; item += npairs) {
      int l, ob, t0, t1; decode(item, l, ob, t0, t1);
      const int o = ob * 256 + (int)rank * 128 + row;
      float bsum = 0.f;
      // column `row` of the D stage: slab row>>6, element row&63 of each of the 64 point-lines
      const int bslab = row >> 6, bc8 = (row & 63) >> 3, be = row & 7;
      for (int t = t0; t < t1; ++t)
        for (int qd = 0; qd < TILE_M / WG_KSTAGE; ++qd) {
          mbar_wait(bar_full(s), ph);
          const uint8_t *st = gS + s * WG_STAGE_BYTES + bslab * WG_SLAB_STAGE;
#pragma unroll 8
          for (int r = 0; r < WG_KSTAGE; ++r) {
            const uint16_t v = *reinterpret_cast<const uint16_t *>(st + sw128_chunk_off(r, bc8) + be * 2);
            bsum += __half2float(__ushort_as_half(v));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty(s));
          if (++s == WG_NSTAGE) { s = 0; ph ^= 1; }
        }
      atomicAdd(p.gB[l] + o, bsum * inv_s);
      // ---- flush dW_l[o, :]
      mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
      tcgen05_fence_after();
      const int ncols = l == 0 ? 128 : D;
      const int ld = l == 0 ? 84 : D;
      float *wrow = p.gW[l] + (int64_t)o * ld;
#pragma unroll 1
      for (int g = 0; g < ncols / 32; ++g) {
        uint32_t acc[32];
        tmem_ld32(tm_row + g * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          int col = g * 32 + i;
          if (l == 0) {
            if (col >= 88) continue;          // zero padding of the encoder image
            if (col >= 84) col -= 84;         // residual columns fold back onto the raw-coordinate weights
          }
          red_add_v4(wrow + col, __uint_as_float(acc[i]) * inv_s, __uint_as_float(acc[i + 1]) * inv_s,
                     __uint_as_float(acc[i + 2]) * inv_s, __uint_as_float(acc[i + 3]) * inv_s);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(accfree_addr);
        else mbar_arrive_remote_relaxed(accfree_addr);
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

// ------------------------------------------------------------------------------------------ output layer grads
// gW_out[o,:] += sum_p g[p,o] h_7[p,:] ; gb_out[o] += sum_p g[p,o].  HBM-bound stream over the saved h_7 images
// (1 KB per point).  256 threads = 4 row groups x 64 column chunks: a warp reads 4 full 128 B image rows per
// instruction, 8 rows (independent 16 B loads) in flight per thread; partial sums stay in registers over all the
// tiles of the CTA, are combined across the row groups in shared memory and leave as one atomic per output.
constexpr int OW_THREADS = 256;
constexpr int OW_ROWS = 8;       // rows (independent 16 B loads) in flight per thread (16 measured slower)
__global__ void __launch_bounds__(OW_THREADS) out_wgrad_bf16_kernel(const float2 *__restrict__ g, int64_t M, int num_tiles,
                                                                    const uint8_t *__restrict__ save_h,
                                                                    float *__restrict__ gW, float *__restrict__ gB) {
  __shared__ float red[4][2 * D + 2];
  const int rg = threadIdx.x >> 6, cg = threadIdx.x & 63;      // rows [32 rg, 32 rg + 32) ; columns [8 cg, 8 cg + 8)
  const int slab = cg >> 3, c8 = cg & 7;
  float a0[8], a1[8], s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const uint8_t *h7 = save_h + ((int64_t)t * NH + (NH - 1)) * A_BYTES + slab * SLAB_BYTES;
#pragma unroll 1
    for (int r0 = rg * 32; r0 < rg * 32 + 32; r0 += OW_ROWS) {
      uint4 hv[OW_ROWS];
      float2 gg[OW_ROWS];
#pragma unroll
      for (int i = 0; i < OW_ROWS; ++i) {
        const int r = r0 + i;
        const int64_t m = (int64_t)t * TILE_M + r;
        hv[i] = __ldcs(reinterpret_cast<const uint4 *>(h7 + sw128_chunk_off(r, c8)));
        gg[i] = m < M ? __ldg(g + m) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < OW_ROWS; ++i) {
        const float h[8] = {h_lo(hv[i].x), h_hi(hv[i].x), h_lo(hv[i].y), h_hi(hv[i].y),
                            h_lo(hv[i].z), h_hi(hv[i].z), h_lo(hv[i].w), h_hi(hv[i].w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] = fmaf(gg[i].x, h[j], a0[j]); a1[j] = fmaf(gg[i].y, h[j], a1[j]); }
        if (cg == 0) { s0 += gg[i].x; s1 += gg[i].y; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[rg][cg * 8 + j] = a0[j]; red[rg][D + cg * 8 + j] = a1[j]; }
  if (cg == 0) { red[rg][2 * D] = s0; red[rg][2 * D + 1] = s1; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D + 2; i += OW_THREADS) {
    const float v = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
    atomicAdd(i < 2 * D ? gW + i : gB + (i - 2 * D), v);
  }
}

}  // namespace bf
}  // namespace snf

using namespace snf;

int snf_bf16_pack_wt(const float *const *W, void *packed, cudaStream_t st) {
  const int64_t chunks = (int64_t)bf::WT_BLOCKS * (bf::WBLK_BYTES / 16);
  bf::pack_wt_kernel<<<(unsigned)ceil_div64(chunks, 256), 256, 0, st>>>(
      W[1], W[2], W[3], W[4], W[5], W[6], W[7], reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_WT_OFF),
      reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_WTLO_OFF));
  count_launch();
  return launch_status();
}

// Per-kernel CUDA-event timing of the backward (bench.py's roofline break-down): off by default; when armed, every call
// records into its own event set (no host sync on the hot path), the sums are formed when the getter is called.
constexpr int BWD_EV_SETS = 256;
static bool g_time_bwd = false;
static cudaEvent_t g_bwd_ev[BWD_EV_SETS][4];
static bool g_bwd_ev_made = false;
static int g_bwd_calls = 0;
extern "C" int snf_debug_time_backward(int on) {
  g_time_bwd = on != 0;
  if (g_time_bwd && !g_bwd_ev_made) {
    for (int s = 0; s < BWD_EV_SETS; ++s)
      for (int i = 0; i < 4; ++i) cudaEventCreate(&g_bwd_ev[s][i]);
    g_bwd_ev_made = true;
  }
  g_bwd_calls = 0;
  return 0;
}
// out[0..2] = summed milliseconds of the dgrad chain, the wgrad and the output-layer gradient kernel over the (at most
// BWD_EV_SETS most recent) timed calls; returns how many calls were summed
extern "C" int snf_debug_backward_ms(double *out) {
  for (int i = 0; i < 3; ++i) out[i] = 0.0;
  const int n = g_bwd_calls < BWD_EV_SETS ? g_bwd_calls : BWD_EV_SETS;
  if (!g_bwd_ev_made || n == 0) return 0;
  cudaDeviceSynchronize();
  for (int s = 0; s < n; ++s)
    for (int i = 0; i < 3; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, g_bwd_ev[s][i], g_bwd_ev[s][i + 1]); out[i] += ms; }
  return n;
}
#define BWD_EV(i) do { if (g_time_bwd) cudaEventRecord(g_bwd_ev[g_bwd_calls % BWD_EV_SETS][i], st); } while (0)

// SNF_DEBUG_SYNC=1: synchronise after every backward kernel and name the one that failed (debug aid only)
static int debug_sync(const char *what, cudaStream_t st) {
  static int on = -1;
  if (on < 0) { const char *e = getenv("SNF_DEBUG_SYNC"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on) return 0;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { fprintf(stderr, "[sunerf_b200] %s failed: %s\n", what, cudaGetErrorString(e)); return (int)e; }
  fprintf(stderr, "[sunerf_b200] %s ok\n", what);
  return 0;
}

int snf_bf16_set_attributes_bwd() {
  cudaError_t e = cudaFuncSetAttribute(bf::mlp_dgrad_bf16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::fw::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(bf::mlp_dgrad_bf16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::fw::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(bf::mlp_wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::WG_SMEM_BYTES);
  return (int)e;
}

int snf_bf16_backward(const float *grad_out, int64_t M, const void *packed, const bf::Bf16Ws &w, float *const *gW,
                      float *const *gB, int num_sms, int wsplit, cudaStream_t st) {
  int num_tiles = (int)((M + bf::TILE_M - 1) / bf::TILE_M);
  num_tiles = (num_tiles + 1) / 2 * 2;   // CTA pairs; the workspace is sized for the padding tile
  // gradients are accumulated with atomics: clear them first (ABI: overwritten).  The trainer hands views of one flat
  // buffer, so adjacent (or alignment-padded) ranges are merged: normally a single memset instead of 18.
  {
    const int64_t wsz[9] = {512 * 84, 512 * 512, 512 * 512, 512 * 512, 512 * 512, 512 * 512, 512 * 512, 512 * 512, 2 * 512};
    uint8_t *lo = nullptr, *hi = nullptr;
    auto flush = [&]() { if (lo != nullptr) cudaMemsetAsync(lo, 0, (size_t)(hi - lo), st); lo = hi = nullptr; };
    auto add = [&](float *ptr, int64_t n) {
      uint8_t *a = reinterpret_cast<uint8_t *>(ptr), *b = a + n * 4;
      if (lo != nullptr && a >= hi && a - hi < 16) { hi = b; return; }   // contiguous up to the 16-byte segment padding
      flush();
      lo = a; hi = b;
    };
    for (int l = 0; l <= bf::NH; ++l) { add(gW[l], wsz[l]); add(gB[l], l < bf::NH ? 512 : 2); }
    flush();
  }
  // max |g| of the batch -> the power-of-two scale of the fp16 gradient images (GradScale)
  cudaMemsetAsync(w.scale, 0, sizeof(bf::GradScale), st);
  {
    int64_t ablocks = ceil_div64(M, 256);
    if (ablocks > 2 * num_sms) ablocks = 2 * num_sms;
    bf::absmax_kernel<<<(unsigned)ablocks, 256, 0, st>>>(reinterpret_cast<const float2 *>(grad_out), M, &w.scale->gmax_bits);
  }
  bf::DgradParams dp{};
  dp.g = reinterpret_cast<const float2 *>(grad_out);
  dp.M = M; dp.num_tiles = num_tiles;
  dp.packed = reinterpret_cast<const uint8_t *>(packed);
  dp.save_pre = w.pre; dp.save_d = w.d; dp.scale = w.scale;
  int grid = num_tiles < num_sms ? num_tiles : num_sms;
  grid &= ~1;
  BWD_EV(0);
  if (wsplit == 2) bf::mlp_dgrad_bf16_kernel<2><<<grid, bf::NTHREADS, bf::fw::SMEM_BYTES, st>>>(dp);
  else bf::mlp_dgrad_bf16_kernel<1><<<grid, bf::NTHREADS, bf::fw::SMEM_BYTES, st>>>(dp);
  BWD_EV(1);
  if (int e = debug_sync("mlp_dgrad_bf16_kernel", st)) return e;

  bf::WgradParams wp{};
  wp.save_d = w.d; wp.save_h = w.h; wp.save_enc = w.enc; wp.scale = w.scale;
  wp.num_tiles = num_tiles;
  // items = tile ranges x 8 layers x 2 o-blocks, about 7 per CTA pair
  const int npairs = num_sms / 2;
  int tpi = (num_tiles * 16 + 7 * npairs - 1) / (7 * npairs);
  if (tpi < 4) tpi = 4;
  wp.tiles_per_item = tpi;
  wp.num_items = ((num_tiles + tpi - 1) / tpi) * 16;
  for (int l = 0; l < bf::NH; ++l) { wp.gW[l] = gW[l]; wp.gB[l] = gB[l]; }
  int wgrid = wp.num_items < npairs ? wp.num_items * 2 : npairs * 2;
  bf::mlp_wgrad_bf16_kernel<<<wgrid, bf::WG_THREADS, bf::WG_SMEM_BYTES, st>>>(wp);
  BWD_EV(2);
  if (int e = debug_sync("mlp_wgrad_bf16_kernel", st)) return e;

  const int ogrid = num_tiles < 4 * num_sms ? num_tiles : 4 * num_sms;   // latency-bound stream: fewer CTAs measured slower
  bf::out_wgrad_bf16_kernel<<<ogrid, bf::OW_THREADS, 0, st>>>(dp.g, M, num_tiles, w.h, gW[bf::NH], gB[bf::NH]);
  BWD_EV(3);
  if (g_time_bwd) ++g_bwd_calls;
  count_launch(4);
  return launch_status();
}
