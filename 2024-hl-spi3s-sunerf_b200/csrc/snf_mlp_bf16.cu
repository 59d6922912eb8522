// bf16 tensor-core field network for sm_100a: positional encoding + 8 x (Linear(512)+sin) + Linear(2) fused in
// ONE persistent, warp-specialised kernel per pass.  Replaces PositionalEncoding.forward / NeRF.forward /
// NeRF_DT.forward (sunerf/model/model.py:123-132, 44-57, 169-187) in "bf16-MLP mode" (BASELINE.json: 1e-2).
//
// Per CTA: one tile of 128 points.  The activations never leave the SM between layers:
//   warp 0      TMA producer : streams the pre-packed bf16 weights (UMMA K-major SWIZZLE_128B image) from L2 with
//                              32 KB cp.async.bulk copies through a 3-stage mbarrier ring
//   warp 1      MMA issuer   : one thread issues tcgen05.mma (M=128, N=256, K=16) per k-step; A = activations in
//                              shared memory (128 KB, same swizzled image), D = 128x512 fp32 in TMEM (all 512 cols)
//   warps 2-5   epilogue     : tcgen05.ld the accumulator, +bias, sin, bf16, st.shared back into the A image for
//                              the next layer; layer 0's A image is the sin/cos encoding computed in place;
//                              the 512->2 output layer is a register dot product fused into the last epilogue
// In training mode every layer's activation image is written to HBM with one TMA bulk store per slab and the
// cosines (the derivative of the sine) with direct 16 B stores, in the same tile-image layout, for the backward.
#include "snf_common.cuh"
#include "snf_tcgen05.cuh"

namespace snf {
namespace bf {
using namespace tc;

constexpr int TILE_M = 128;
constexpr int D = 512;
constexpr int NH = 8;
constexpr int K0 = 96;                       // layer-0 K: 84 features + 4 bf16 residuals of x + 8 zero columns
constexpr int SLAB_BYTES = TILE_M * 128;     // one K-slab (64 bf16) of a 128-row A image: 16 KB
constexpr int A_BYTES = 8 * SLAB_BYTES;      // 128 KB
constexpr int WBLK_ROWS = 256;               // weight block: 256 output features x 64 k
constexpr int WBLK_BYTES = WBLK_ROWS * 128;  // 32 KB
constexpr int NSTAGE = 3;
constexpr int NTHREADS = 192;
constexpr int SMEM_BYTES = A_BYTES + NSTAGE * WBLK_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__host__ __device__ constexpr int layer_slabs(int l) { return l == 0 ? 2 : 8; }
__host__ __device__ constexpr int layer_blocks(int l) { return 2 * layer_slabs(l); }
constexpr int TOTAL_BLOCKS = 4 + 7 * 16;   // 116 weight blocks of 32 KB
// packed buffer: [TOTAL_BLOCKS x 32 KB bf16 images][bias 8x512 f32][W_out 2x512 f32][b_out 2 f32 (+2 pad)]
constexpr int64_t PACK_W_BYTES = (int64_t)TOTAL_BLOCKS * WBLK_BYTES;
constexpr int64_t PACK_BYTES = PACK_W_BYTES + (NH * D + 2 * D + 4) * 4;

// ------------------------------------------------------------------------------------------ weight packing
// one thread per 16-byte chunk of the packed image
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *w0, const float *w1, const float *w2,
                                                           const float *w3, const float *w4, const float *w5,
                                                           const float *w6, const float *w7, uint4 *__restrict__ dst) {
  const float *W[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)TOTAL_BLOCKS * (WBLK_BYTES / 16)) return;
  const int blk = (int)(idx / (WBLK_BYTES / 16));
  const int within = (int)(idx % (WBLK_BYTES / 16));
  // block order = consumption order: layer, n-half, k-slab
  int l, nh, ks;
  if (blk < 4) { l = 0; nh = blk >> 1; ks = blk & 1; }
  else { const int b = blk - 4; l = 1 + b / 16; nh = (b % 16) >> 3; ks = b & 7; }
  const int r = within >> 3;                 // row inside the block (0..255)
  const int pos = within & 7;                // stored chunk position inside the 128 B row
  const int c8 = pos ^ (r & 7);              // logical chunk (SWIZZLE_128B)
  const int n = nh * WBLK_ROWS + r;
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
  int num_tiles;
  const uint8_t *packed;  // PACK_BYTES
  float2 *out;            // [M]
  float off0, off1;
  uint8_t *save_enc;      // train: [tiles][2 slabs][16 KB] bf16 image, else null
  uint8_t *save_h;        // train: [tiles][8][128 KB]
  uint8_t *save_c;        // train: [tiles][8][128 KB]
};

__global__ void __launch_bounds__(NTHREADS, 1) mlp_fwd_bf16_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sW = base + A_BYTES;
  const uint32_t sBar = sW + NSTAGE * WBLK_BYTES;
  uint8_t *gA = smem_raw + (base - smem_u32(smem_raw));   // generic pointer to the A image
  // barriers: full[NSTAGE], empty[NSTAGE], acc_full, a_ready ; then the TMEM base slot
  auto bar_full = [&](int s) { return sBar + 8u * s; };
  auto bar_empty = [&](int s) { return sBar + 8u * (NSTAGE + s); };
  const uint32_t bar_acc = sBar + 8u * (2 * NSTAGE), bar_aready = sBar + 8u * (2 * NSTAGE + 1);
  const uint32_t tmem_slot = sBar + 8u * (2 * NSTAGE + 2);
  volatile uint32_t *tmem_slot_g = reinterpret_cast<volatile uint32_t *>(gA + A_BYTES + NSTAGE * WBLK_BYTES + 8 * (2 * NSTAGE + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_aready, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot_g;

  const float *bias_all = reinterpret_cast<const float *>(p.packed + PACK_W_BYTES);
  const float *w_out = bias_all + NH * D;
  const float *b_out = w_out + 2 * D;

  if (warp == 0) {
    // =========================== TMA producer
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int blk = 0; blk < TOTAL_BLOCKS; ++blk) {
          mbar_wait(bar_empty(s), ph ^ 1);
          mbar_arrive_expect_tx(bar_full(s), WBLK_BYTES);
          bulk_g2s(sW + s * WBLK_BYTES, p.packed + (int64_t)blk * WBLK_BYTES, WBLK_BYTES, bar_full(s));
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(128, 256);
      int s = 0; uint32_t ph = 0, ph_a = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < NH; ++l) {
          mbar_wait(bar_aready, ph_a); ph_a ^= 1;     // this layer's A image is complete, TMEM is drained
          tcgen05_fence_after();
          const int nslab = layer_slabs(l);
          for (int nh = 0; nh < 2; ++nh) {
            for (int ks = 0; ks < nslab; ++ks) {
              mbar_wait(bar_full(s), ph);
              tcgen05_fence_after();
              const int ksteps = (l == 0 && ks == 1) ? (K0 - 64) / 16 : 4;
#pragma unroll 4
              for (int k4 = 0; k4 < ksteps; ++k4) {
                const uint64_t ad = smem_desc(sA + ks * SLAB_BYTES + k4 * 32, 16, 1024);
                const uint64_t bd = smem_desc(sW + s * WBLK_BYTES + k4 * 32, 16, 1024);
                mma_ss(tmem + nh * 256, ad, bd, idesc, (ks | k4) != 0);
              }
              mma_commit(bar_empty(s));               // frees the weight stage when these MMAs have read it
              if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
          }
          mma_commit(bar_acc);                        // whole layer accumulated
        }
      }
    }
  } else {
    // =========================== epilogue warps (128 threads, thread == row)
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;                  // 0..127
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16);
    const bool train = p.save_h != nullptr;
    uint32_t ph_acc = 0;
    // chunks 4..7 of slab 1 are never read by the layer-0 MMAs but are part of the saved encoder image
    for (int c8 = 4; c8 < 8; ++c8)
      *reinterpret_cast<uint4 *>(gA + SLAB_BYTES + sw128_chunk_off(row, c8)) = make_uint4(0, 0, 0, 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t m = (int64_t)tile * TILE_M + row;
      // ---- layer-0 operand: positional encoding of this row, written straight into the A image
      {
        if (train) { if (et == 0) bulk_wait_read_all(); named_bar_sync(1, 128); }
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M) xv = p.x[m];
        const float xc[4] = {xv.x, xv.y, xv.z, xv.w};
        float feat[K0];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          feat[c] = xc[c];
          feat[84 + c] = xc[c] - __bfloat162float(__float2bfloat16_rn(xc[c]));   // what bf16 drops from x
        }
#pragma unroll
        for (int f = 0; f < 10; ++f)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float sv, cv;
            sincosf(xc[c] * (float)(1 << f) * 0.5f, &sv, &cv);   // exact scalings, accurate range reduction
            feat[4 + f * 4 + c] = sv;
            feat[44 + f * 4 + c] = cv;
          }
#pragma unroll
        for (int k = 88; k < K0; ++k) feat[k] = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < K0 / 8; ++c8) {
          uint4 o;
          o.x = pack_bf16x2(feat[c8 * 8 + 0], feat[c8 * 8 + 1]); o.y = pack_bf16x2(feat[c8 * 8 + 2], feat[c8 * 8 + 3]);
          o.z = pack_bf16x2(feat[c8 * 8 + 4], feat[c8 * 8 + 5]); o.w = pack_bf16x2(feat[c8 * 8 + 6], feat[c8 * 8 + 7]);
          *reinterpret_cast<uint4 *>(gA + (c8 >> 3) * SLAB_BYTES + sw128_chunk_off(row, c8 & 7)) = o;
        }
        fence_proxy_async_smem();
        if (train) {
          named_bar_sync(1, 128);
          if (et == 0) {
            bulk_s2g(p.save_enc + (int64_t)tile * 2 * SLAB_BYTES, sA, 2 * SLAB_BYTES);
            bulk_commit();
          }
        }
        tcgen05_fence_before();
        mbar_arrive(bar_aready);
      }
      // ---- layers
      for (int l = 0; l < NH; ++l) {
        mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
        tcgen05_fence_after();
        const bool last = (l == NH - 1);
        const bool write_a = !last || train;
        if (train) { if (et == 0) bulk_wait_read_all(); named_bar_sync(1, 128); }
        const float *bias = bias_all + l * D;
        uint8_t *csave = train ? p.save_c + ((int64_t)tile * NH + l) * A_BYTES : nullptr;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll 1
        for (int g = 0; g < D / 32; ++g) {
          uint32_t acc[32];
          tmem_ld32(tm_row + g * 32, acc);
          tmem_ld_wait();
          float hv[32];
          uint32_t cpk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + g * 32 + i));
            const float v0 = __uint_as_float(acc[i]) + b.x, v1 = __uint_as_float(acc[i + 1]) + b.y;
            const float v2 = __uint_as_float(acc[i + 2]) + b.z, v3 = __uint_as_float(acc[i + 3]) + b.w;
            hv[i] = __sinf(v0); hv[i + 1] = __sinf(v1); hv[i + 2] = __sinf(v2); hv[i + 3] = __sinf(v3);
            if (train) {
              cpk[i / 2] = pack_bf16x2(__cosf(v0), __cosf(v1));
              cpk[i / 2 + 1] = pack_bf16x2(__cosf(v2), __cosf(v3));
            }
          }
          if (last) {   // fused output layer: out = W_out h + b_out
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 wa = __ldg(reinterpret_cast<const float4 *>(w_out + g * 32 + i));
              const float4 wb = __ldg(reinterpret_cast<const float4 *>(w_out + D + g * 32 + i));
              o0 += hv[i] * wa.x + hv[i + 1] * wa.y + hv[i + 2] * wa.z + hv[i + 3] * wa.w;
              o1 += hv[i] * wb.x + hv[i + 1] * wb.y + hv[i + 2] * wb.z + hv[i + 3] * wb.w;
            }
          }
          const int slab = g >> 1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int c8 = (g & 1) * 4 + c;
            const uint32_t off = slab * SLAB_BYTES + sw128_chunk_off(row, c8);
            if (write_a) {
              uint4 o;
              o.x = pack_bf16x2(hv[c * 8 + 0], hv[c * 8 + 1]); o.y = pack_bf16x2(hv[c * 8 + 2], hv[c * 8 + 3]);
              o.z = pack_bf16x2(hv[c * 8 + 4], hv[c * 8 + 5]); o.w = pack_bf16x2(hv[c * 8 + 6], hv[c * 8 + 7]);
              *reinterpret_cast<uint4 *>(gA + off) = o;
            }
            if (train)
              *reinterpret_cast<uint4 *>(csave + off) = make_uint4(cpk[c * 4], cpk[c * 4 + 1], cpk[c * 4 + 2], cpk[c * 4 + 3]);
          }
        }
        if (write_a) fence_proxy_async_smem();
        if (train) {
          named_bar_sync(1, 128);
          if (et == 0) {
            uint8_t *dst = p.save_h + ((int64_t)tile * NH + l) * A_BYTES;
#pragma unroll 1
            for (int sl = 0; sl < 8; ++sl) bulk_s2g(dst + sl * SLAB_BYTES, sA + sl * SLAB_BYTES, SLAB_BYTES);
            bulk_commit();
          }
        }
        if (last) {
          if (m < p.M) p.out[m] = make_float2(o0 + __ldg(b_out) + p.off0, o1 + __ldg(b_out + 1) + p.off1);
        } else {
          tcgen05_fence_before();
          mbar_arrive(bar_aready);
        }
      }
    }
    if (train && et == 0) bulk_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

struct Bf16Ws {
  uint8_t *enc, *h, *c, *d;
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
    w.c = p + off; off += tiles * NH * (int64_t)A_BYTES;
    w.d = p + off; off += tiles * NH * (int64_t)A_BYTES;   // dpre_l images written by the backward
  }
  w.bytes = off > 0 ? off : 256;
  return w;
}

}  // namespace bf
}  // namespace snf

using namespace snf;

int64_t snf_mlp_bf16_ws_bytes(int64_t M, int train) { return bf::bf16_layout(nullptr, M, train).bytes; }

// snf_mlp_bf16_bwd.cu
int64_t snf_bf16_pack_total_bytes();
int snf_bf16_pack_wt(const float *const *W, void *packed, cudaStream_t st);
int snf_bf16_backward(const float *grad_out, int64_t M, const void *packed, uint8_t *enc, uint8_t *h, uint8_t *c, uint8_t *d,
                      float *const *gW, float *const *gB, int num_sms, cudaStream_t st);

extern "C" int64_t snf_mlp_pack_bytes(void) { return snf_bf16_pack_total_bytes(); }

extern "C" int snf_mlp_pack_bf16(const float *const *W, const float *const *B, void *packed, void *stream) {
  SNF_CHECK_PTR(W); SNF_CHECK_PTR(B); SNF_CHECK_PTR(packed); SNF_CHECK_ALIGN(packed, 1024);
  for (int l = 0; l <= bf::NH; ++l) { SNF_CHECK_PTR(W[l]); SNF_CHECK_PTR(B[l]); }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = (int64_t)bf::TOTAL_BLOCKS * (bf::WBLK_BYTES / 16);
  bf::pack_weights_kernel<<<(unsigned)ceil_div64(chunks, 256), 256, 0, st>>>(W[0], W[1], W[2], W[3], W[4], W[5], W[6], W[7],
                                                                          reinterpret_cast<uint4 *>(packed));
  const int nsmall = bf::NH * bf::D + 2 * bf::D + 2;
  bf::pack_small_kernel<<<(nsmall + 255) / 256, 256, 0, st>>>(B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], W[8], B[8],
                                                             reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(packed) + bf::PACK_W_BYTES));
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
    cudaError_t e = cudaFuncSetAttribute(bf::mlp_fwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  bf::FwdParams p{};
  p.x = reinterpret_cast<const float4 *>(x);
  p.M = M;
  p.num_tiles = (int)((M + bf::TILE_M - 1) / bf::TILE_M);
  p.packed = reinterpret_cast<const uint8_t *>(packed);
  p.out = reinterpret_cast<float2 *>(out);
  p.off0 = off0; p.off1 = off1;
  if (train) {
    bf::Bf16Ws w = bf::bf16_layout(ws, M, 1);
    p.save_enc = w.enc; p.save_h = w.h; p.save_c = w.c;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  bf::mlp_fwd_bf16_kernel<<<grid, bf::NTHREADS, bf::SMEM_BYTES, (cudaStream_t)stream>>>(p);
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
  return snf_bf16_backward(grad_out, M, packed, w.enc, w.h, w.c, w.d, gW, gB, num_sms(), (cudaStream_t)stream);
}
