// Constants, layouts and role helpers shared by the 16-bit tensor-core kernels (forward, dgrad chain, wgrad).
// Operand format: fp16 for every MMA operand (weights, activations, dL/dpre), fp32 accumulation.  The path keeps the
// name the task gives it ("bf16-MLP mode"); round 1 ran bf16 operands, whose 8-bit significand on the WEIGHTS alone costs
// 5.7e-3 of per-parameter gradient accuracy against the 1e-3 gate (tools/micro/bf16_grad_study.py: fp16 weights 6.8e-4).
// tcgen05 kind::f16 rejects A and B of different 16-bit formats (illegal instruction, tools/umma_probe.cu "mixed"), so the
// back-propagated gradients are fp16 too, carried with one power-of-two scale per backward call (GradScale below).
#pragma once
#include "snf_common.cuh"
#include "snf_tcgen05.cuh"

namespace snf {
namespace bf {
using namespace tc;

constexpr int TILE_M = 128;                  // points per CTA tile (= TMEM lanes); a CTA pair covers 256
constexpr int D = 512;                       // hidden width
constexpr int NH = 8;                        // hidden layers
constexpr int K0 = 96;                       // layer-0 K: 84 features + 4 fp16 residuals of x + 8 zero columns
constexpr int SLAB_BYTES = TILE_M * 128;     // one K-slab (64 bf16) of a 128-row image: 16 KB
constexpr int A_BYTES = 8 * SLAB_BYTES;      // 128 KB activation image
constexpr int C_BYTES = TILE_M * D;          // 64 KB: cos(pre) of one tile and layer as int8 (x 127), chunk-major:
                                             // [slab (8)][16-column chunk (4)][row (128)] x 16 B; a warp access covers 512 B

// ---- layer-chain kernels (forward, dgrad): CTA pairs, tcgen05 cta_group::2, M=256 N=256 K=16
constexpr int NCHUNK = 256;                  // output features per MMA / per weight block
constexpr int WBLK_BYTES = NCHUNK * 128;     // weight block: 256 output features x 64 k = 32 KB
constexpr int WHALF_BYTES = WBLK_BYTES / 2;  // each CTA of the pair streams half of every block: 16 KB
#ifndef SNF_EPI_GROUPS
#define SNF_EPI_GROUPS 2                     // per translation unit: the forward uses 4 (16 epilogue warps), the dgrad chain 2
#endif
constexpr int EPI_GROUPS = SNF_EPI_GROUPS;   // epilogue warps per TMEM lane quarter: each owns 64 / EPI_GROUPS columns of a step
constexpr int CPT = 64 / EPI_GROUPS;         // accumulator columns per thread and step (a step = one 64-column k-slab)
constexpr int CHUNKS = CPT / 8;              // 16-byte chunks of the A image per thread and step
constexpr int N_EPI_WARPS = 4 * EPI_GROUPS;
constexpr int N_EPI = N_EPI_WARPS * 32;      // epilogue threads: (row, column group)
constexpr int EPI_WARP0 = 4;                 // warpgroup 0: warp 0 TMA producer, warp 1 MMA issuer / peer relay, warp 2 TMA store warp (training, dgrad);
constexpr int NTHREADS = 128 + N_EPI;        // the following warpgroups: epilogue.  Registers are re-balanced with setmaxnreg:
constexpr int REGS_CTRL = 40, REGS_EPI = EPI_GROUPS == 2 ? 232 : 104;   // 128 x CTRL + N_EPI x EPI must fit the launch allocation (NTHREADS x 168 or 96)
static_assert(EPI_GROUPS == 2 || EPI_GROUPS == 4, "CPT must be 32 or 16 (tcgen05.ld x32 / x16)");
constexpr int BIAS_BYTES = D * 4;

constexpr int FWD_BLOCKS = 2 * 2 + 7 * 16;   // forward weight blocks: layer 0 (2 n-halves x 2 k-slabs) + 7 x (2 x 8)
constexpr int WT_BLOCKS = 7 * 16;            // W^T blocks for the dgrad chain (layers 1..7)
// packed buffer: [FWD_BLOCKS x 32 KB][bias 8x512 f32][W_out 2x512 f32][b_out 2 f32 + pad] | [WT_BLOCKS x 32 KB] |
//                [FWD_BLOCKS x 32 KB: fp16(W - fp16(W)), the low halves for the split-precision forward] |
//                [WT_BLOCKS x 32 KB: low halves of the W^T blocks for the split-precision dgrad chain]
constexpr int64_t PACK_W_BYTES = (int64_t)FWD_BLOCKS * WBLK_BYTES;
constexpr int64_t PACK_BIAS_OFF = PACK_W_BYTES;
constexpr int64_t PACK_WOUT_OFF = PACK_BIAS_OFF + NH * D * 4;
constexpr int64_t PACK_BOUT_OFF = PACK_WOUT_OFF + 2 * D * 4;
constexpr int WOUT_BYTES = 2 * D * 4;        // 4 KB, rides through the weight ring as a pseudo-block
constexpr int64_t PACK_WT_OFF = (PACK_BOUT_OFF + 16 + 1023) / 1024 * 1024;
constexpr int64_t PACK_LO_OFF = PACK_WT_OFF + (int64_t)WT_BLOCKS * WBLK_BYTES;
constexpr int64_t PACK_WTLO_OFF = PACK_LO_OFF + (int64_t)FWD_BLOCKS * WBLK_BYTES;
constexpr int64_t PACK_TOTAL_BYTES = PACK_WTLO_OFF + (int64_t)WT_BLOCKS * WBLK_BYTES;

constexpr int FMT = FMT_F16;                 // operand format of every MMA of the field network


__device__ __forceinline__ float h_lo(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xFFFFu))); }
__device__ __forceinline__ float h_hi(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }

// cos(pre) travels from the forward to the dgrad chain as ONE byte: bit 7 = (cos < 0), bits 0..6 = q = rint(127 t) with
// t = sqrt(1 - |cos|) (127.014 code
// units per t, COSQ_T); decode |cos| = 1 - (q / 127.014)^2.  The resolution is finest where the cosines pile up - next to +-1,
// i.e. small pre-activations - and 2/254 at the zero crossings.  Round 1 stored rint(127 cos): every cosine above
// 1 - 1/254 came back as exactly 1, a one-sided error that does not average out over points and cost 3.4e-3 of
// per-parameter gradient accuracy by itself; this code costs 1.1e-4 (tools/micro/bf16_grad_study.py, 'i8half').
// With the half-angle values s2 = sin(pre/2), c2 = cos(pre/2) the forward already holds: t = sqrt(2) min(|s2|, |c2|),
// cos < 0 <=> |s2| > |c2|, and sin(pre) = 2 s2 c2 - two MUFU per element as before.
// Encode: t' + 1.5 * 2^23 leaves the integer in the low mantissa byte.
constexpr float COSQ_MAGIC = 12582912.f;
constexpr float COSQ_SCALE = 179.625f;       // code units per min(|s|, |c|): 127 sqrt(2) rounded to an fp16 value (0.011 % above)
constexpr float COSQ_T = 127.0140556f;       // = COSQ_SCALE / sqrt(2): code units per t = sqrt(1 - |cos|)
constexpr float COSQ_INV_T2 = 6.1986403e-5f;  // 1 / COSQ_T^2
__device__ __forceinline__ uint32_t cosq_enc(float s2, float c2) {
  const float as = fabsf(s2), ac = fabsf(c2);
  const uint32_t q = __float_as_uint(fmaf(fminf(as, ac), COSQ_SCALE, COSQ_MAGIC));
  return as > ac ? (q | 0x80u) : q;
}
// Two codes at once from fp16 pairs (s, c) = (sin, cos)(pre / 2): byte 0 and byte 2 of the result hold the codes of the low
// and high element.  min(|s|, |c|) 127 sqrt(2) + 1024 is an fp16 integer 1024 + q (spacing 1 there: the add rounds to
// nearest), i.e. the halves 0x64qq.
__device__ __forceinline__ uint32_t cosq_enc2(__half2 s, __half2 c) {
  const __half2 as = __habs2(s), ac = __habs2(c);
  const __half2 qh = __hfma2(__hmin2(as, ac), __float2half2_rn(COSQ_SCALE), __float2half2_rn(1024.f));
  const uint32_t neg = __hgt2_mask(as, ac);                      // 0xFFFF per element with cos(pre) < 0
  return (*reinterpret_cast<const uint32_t *>(&qh) & 0x00FF00FFu) | (neg & 0x00800080u);
}
__device__ __forceinline__ uint32_t cosq_pack4(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3) {   // low bytes -> 4 codes
  return __byte_perm(__byte_perm(t0, t1, 0x0040), __byte_perm(t2, t3, 0x0040), 0x5410);
}
// Decode code B (0..3) of the word w in fp32: the byte lands in the low mantissa bits of 2^23, the subtraction is exact.
// (An fp16 decode was measured first: fp16 cannot tell 1 - q^2/T^2 from 1 for q <= 2 and rounds every cosine above 0.9998
// to exactly 1 - the round-1 pile-up again, 16 times smaller; together with an fp16 x fp16 product it cost 4.5e-4 of
// per-parameter gradient accuracy, against 1.1e-4 for this form.  tools/micro/bf16_grad_study.py, 'x3 fp16 decode'.)
template <int B> __device__ __forceinline__ float cosq_dec_abs(uint32_t w) {   // |cos| of code B
  const uint32_t f = __byte_perm(w & 0x7F7F7F7Fu, 0x4B000000u, 0x7650 + B);
  const float q = __uint_as_float(f) - 8388608.f;
  return fmaf(q * q, -COSQ_INV_T2, 1.f);
}
// the signs of codes 2 PR, 2 PR + 1 as the sign bits of an fp16 pair: XOR it into the packed products
template <int PR> __device__ __forceinline__ uint32_t cosq_sign2(uint32_t w) {
  return __byte_perm(w, 0u, PR == 0 ? 0x1404 : 0x3424) & 0x80008000u;
}

// One power-of-two scale per backward call keeps the fp16 dL/dpre images in range: S = 2^-ceil(log2(bound)) with
// bound = max_p max(|g0|, |g1|) * max_j (|W_out[0,j]| + |W_out[1,j]|) >= max |dL/dh_7|, so the first image peaks in
// (0.5, 1] and the chain may grow 2^16-fold before it overflows; elements 2^-24 below the peak flush to zero (their
// share of a gradient sum is below fp32 resolution).  The whole backward is linear in g, so the wgrad un-scales exactly
// when it flushes its fp32 accumulators.  ws.scale = {absmax bits of g (uint), S, 1/S, pad}.
struct GradScale { uint32_t gmax_bits; float S, invS, pad; };
__device__ __forceinline__ float grad_scale_from(float gmax, float wmax) {
  const float bound = gmax * wmax;
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;            // all-zero, Inf or NaN gradients: nothing to protect
  int e;
  frexpf(bound, &e);                                               // bound = m 2^e, m in [0.5, 1)
  e = e > 126 ? 126 : e < -126 ? -126 : e;
  return ldexpf(1.f, -e);
}

// ---- layer-chain kernels (forward, dgrad chain): shared-memory layout and barrier block
//   [A image 128 KB][weight ring 5 x 16 KB][bias: 2 layers x 2 KB][W_out 4 KB][row partial sums 1 KB][barriers]
namespace fw {
constexpr int NSTAGE = 5;
constexpr int OFF_RING = A_BYTES;
constexpr int OFF_BIAS = OFF_RING + NSTAGE * WHALF_BYTES;
constexpr int OFF_WOUT = OFF_BIAS + 2 * BIAS_BYTES;
constexpr int OFF_OSUM = OFF_WOUT + WOUT_BYTES;
constexpr int OFF_BAR = OFF_OSUM + (EPI_GROUPS - 1) * TILE_M * 8;
constexpr int SMEM_BYTES = OFF_BAR + 512;
static_assert(SMEM_BYTES <= 232448, "layer-chain kernels exceed the 227 KB shared-memory window");
//   full[s] / empty[s] : weight ring, as in Bars
//   acc[h]             : temporal N-half h (output columns [256h, 256h+256)) of the current layer accumulated
//                        (multicast tcgen05.commit -> both CTAs)
//   ready[k] (leader)  : 16 arrivals = 8 epilogue warps x 2 CTAs.  k=0: A slabs 0..3 written (and D half 0 drained);
//                        k=1..4: A slab 3+k written; k=4 also means D half 1 drained
struct Bars {
  uint32_t base;
  __device__ uint32_t full(int s) const { return base + 8u * s; }
  __device__ uint32_t empty(int s) const { return base + 8u * (NSTAGE + s); }
  __device__ uint32_t acc(int h) const { return base + 8u * (2 * NSTAGE + h); }
  __device__ uint32_t ready(int k) const { return base + 8u * (2 * NSTAGE + 2 + k); }
  __device__ uint32_t tmem_slot() const { return base + 8u * (2 * NSTAGE + 7); }
  // training / dgrad, local to each CTA: wrote[slot] = all epilogue warps have written (and fenced) the slab(s) of this
  // event into the A image (slot 0: slabs 0-3 at once or the encoder image, slot k: slab k); afree = the store warp's
  // TMA stores of the previous layer have finished reading the A image
  __device__ uint32_t wrote(int slot) const { return base + 8u * (2 * NSTAGE + 8 + slot); }
  __device__ uint32_t afree() const { return base + 8u * (2 * NSTAGE + 16); }
  // acc1a: the MMAs of half 1 over k-slabs 0..3 are complete (multicast commit): no instruction of this layer reads slabs
  // 0..3 of the A image any more, half 0 of the next operand may be written while half 1 is still being accumulated
  __device__ uint32_t acc1a() const { return base + 8u * (2 * NSTAGE + 17); }
};
constexpr int TMEM_SLOT_OFF = OFF_BAR + 8 * (2 * NSTAGE + 7);
static_assert(8 * (2 * NSTAGE + 18) <= 512, "barrier block");
}  // namespace fw

// saved-image workspace (training): [enc][H = sin(pre)][D = S dL/dpre] as [tile][layer][128 KB] fp16 images,
// [C = cos(pre)] as [tile][layer][64 KB] one-byte codes, and the GradScale block of the backward.
// Sized for an even number of tiles (CTA pairs always process two).
struct Bf16Ws {
  uint8_t *enc, *h, *pre, *d;
  GradScale *scale;
  // split-precision ("x3") forward only: low halves of the encoder image and of the running activations (ping-pong),
  // high halves as ping-pong when nothing is saved (inference), per-(point, N-half, column group) partial outputs
  uint8_t *enc_lo, *h_lo[2], *h_pp[2];
  float2 *part;
  int64_t bytes;
};
// x3: 0 = the fused 16-bit kernels' workspace; 1 = plus what the per-layer split-precision forward needs
inline Bf16Ws bf16_layout(void *base, int64_t M, int train, int x3 = 0) {
  Bf16Ws w{};
  int64_t tiles = (M + TILE_M - 1) / TILE_M;
  tiles = (tiles + 1) / 2 * 2;
  uint8_t *p = reinterpret_cast<uint8_t *>(base);
  int64_t off = 0;
  if (train) {
    w.enc = p + off; off += tiles * 2 * SLAB_BYTES;
    w.h = p + off; off += tiles * NH * (int64_t)A_BYTES;
    w.pre = p + off; off += tiles * NH * (int64_t)C_BYTES;   // quantised cos(pre), see C_BYTES
    w.d = p + off; off += tiles * NH * (int64_t)A_BYTES;
    w.scale = reinterpret_cast<GradScale *>(p + off); off += 256;
  }
  if (x3) {
    if (!train) { w.enc = p + off; off += tiles * 2 * SLAB_BYTES; }
    w.enc_lo = p + off; off += tiles * 2 * SLAB_BYTES;
    for (int i = 0; i < 2; ++i) { w.h_lo[i] = p + off; off += tiles * (int64_t)A_BYTES; }
    if (!train)
      for (int i = 0; i < 2; ++i) { w.h_pp[i] = p + off; off += tiles * (int64_t)A_BYTES; }
    w.part = reinterpret_cast<float2 *>(p + off); off += tiles * TILE_M * 4 * (int64_t)sizeof(float2);
  }
  w.bytes = off > 0 ? off : 256;
  return w;
}

}  // namespace bf
}  // namespace snf
