// fp32 field network (the 1e-5 parity mode): positional encoding + sine MLP forward and backward on the FP32
// SIMT pipes.  Replaces PositionalEncoding.forward (sunerf/model/model.py:123-132), NeRF.forward (:44-57) and
// NeRF_DT.forward (:169-187) and the autograd backward torch derives for them.  The bf16 tensor-core mode that
// carries the throughput lives in snf_mlp_bf16.cu; this one exists because TF32/bf16 cannot hold 1e-5 through
// eight sine layers and an exp head.
#include "snf_common.cuh"

namespace snf {

constexpr int ENC_F = 10;                 // n_freqs (model.py:29)
constexpr int ENC_OUT = 4 * (1 + 2 * ENC_F);   // 84

// ------------------------------------------------------------------------------------------ encoding (a4)
// out[m] = [x(4), sin(x_c * 2^f / 2) (f major, c minor)(40), cos(same)(40)]
__global__ void __launch_bounds__(256) encode_kernel(const float4 *__restrict__ x, int64_t M,
                                                     float *__restrict__ enc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 44) return;
  const int64_t m = idx / 44;
  const int e = (int)(idx - m * 44);
  const float4 p = x[m];
  const float v[4] = {p.x, p.y, p.z, p.w};
  float *row = enc + m * ENC_OUT;
  if (e < 4) { row[e] = v[e]; return; }
  const int f = (e - 4) >> 2, c = (e - 4) & 3;
  const float arg = fdiv(fmul(v[c], (float)(1 << f)), 2.f);   // x * freq / scale_factor, both exact scalings
  float s, co;
  sincosf(arg, &s, &co);
  row[4 + f * 4 + c] = s;
  row[4 + 4 * ENC_F + f * 4 + c] = co;
}

// ------------------------------------------------------------------------------------------ SGEMM
// C[M,N] (+epilogue) = op(A) * op(B).  128x128x16 tiles, 256 threads, 8x8 register tile per thread (as 2x2 blocks
// of 4x4 so shared-memory reads are conflict-free float4s), register-staged double buffering.
//   ALAY 0: A[m*lda + k] (k contiguous)      ALAY 1: A[k*lda + m] (m contiguous)
//   BLAY 0: B[n*ldb + k] (k contiguous)      BLAY 1: B[k*ldb + n] (n contiguous)
//   EPI 0 : v = acc + bias[n]; H = sin(v); Cs = cos(v) (if non-null)          forward layer
//   EPI 1 : D = acc * Cs[m,n]                                                  dgrad through the sine
//   EPI 2 : P[blockIdx.z][m,n] = acc                                           split-K wgrad partial
constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;

template <int ALAY, int BLAY, int EPI>
__global__ void __launch_bounds__(256)
    sgemm_kernel(const float *__restrict__ A, int lda, const float *__restrict__ B, int ldb, int64_t M, int N,
                 int64_t K, int64_t k_per_split, const float *__restrict__ bias, const float *__restrict__ Cin,
                 float *__restrict__ Out, float *__restrict__ Out2, int ldc) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads; thread tile rows {ty*4+i, 64+ty*4+i}, cols likewise

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (ALAY == 0) {   // 128 rows x 16 k: thread -> row tid/4 + 64h, k quad (tid%4)*4
        const int64_t m = m0 + (tid >> 2) + 64 * h;
        const int64_t k = k0 + (tid & 3) * 4;
        ra[h] = (m < M && k < kend) ? *reinterpret_cast<const float4 *>(A + m * lda + k) : make_float4(0, 0, 0, 0);
      } else {           // 16 k x 128 m: thread -> k tid/32 + 8h, m quad (tid%32)*4
        const int64_t k = k0 + (tid >> 5) + 8 * h;
        const int64_t m = m0 + (tid & 31) * 4;
        ra[h] = (k < kend && m < M) ? *reinterpret_cast<const float4 *>(A + k * lda + m) : make_float4(0, 0, 0, 0);
      }
      if (BLAY == 0) {
        const int n = n0 + (tid >> 2) + 64 * h;
        const int64_t k = k0 + (tid & 3) * 4;
        rb[h] = (n < N && k < kend) ? *reinterpret_cast<const float4 *>(B + (int64_t)n * ldb + k) : make_float4(0, 0, 0, 0);
      } else {
        const int64_t k = k0 + (tid >> 5) + 8 * h;
        const int n = n0 + (tid & 31) * 4;
        rb[h] = (k < kend && n < N) ? *reinterpret_cast<const float4 *>(B + k * ldb + n) : make_float4(0, 0, 0, 0);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (ALAY == 0) {
        const int r = (tid >> 2) + 64 * h, kq = (tid & 3) * 4;
        As[buf][kq + 0][r] = ra[h].x; As[buf][kq + 1][r] = ra[h].y; As[buf][kq + 2][r] = ra[h].z; As[buf][kq + 3][r] = ra[h].w;
      } else {
        *reinterpret_cast<float4 *>(&As[buf][(tid >> 5) + 8 * h][(tid & 31) * 4]) = ra[h];
      }
      if (BLAY == 0) {
        const int r = (tid >> 2) + 64 * h, kq = (tid & 3) * 4;
        Bs[buf][kq + 0][r] = rb[h].x; Bs[buf][kq + 1][r] = rb[h].y; Bs[buf][kq + 2][r] = rb[h].z; Bs[buf][kq + 3][r] = rb[h].w;
      } else {
        *reinterpret_cast<float4 *>(&Bs[buf][(tid >> 5) + 8 * h][(tid & 31) * 4]) = rb[h];
      }
    }
  };

  const int64_t nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  if (nk > 0) { load_tiles(kbeg); store_tiles(0); }
  __syncthreads();
  for (int64_t it = 0; it < nk; ++it) {
    const int buf = (int)(it & 1);
    if (it + 1 < nk) load_tiles(kbeg + (it + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (it + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  float *outp = Out;
  if (EPI == 2) outp = Out + (int64_t)blockIdx.z * M * ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + (jh == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= N) continue;   // N % 4 == 0, so a quad is all-in or all-out
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      const int64_t off = m * ldc + n;
      if (EPI == 0) {
        const float4 b = *reinterpret_cast<const float4 *>(bias + n);
        float4 s, c;
        sincosf(v.x + b.x, &s.x, &c.x); sincosf(v.y + b.y, &s.y, &c.y);
        sincosf(v.z + b.z, &s.z, &c.z); sincosf(v.w + b.w, &s.w, &c.w);
        *reinterpret_cast<float4 *>(outp + off) = s;
        if (Out2 != nullptr) *reinterpret_cast<float4 *>(Out2 + off) = c;
      } else if (EPI == 1) {
        const float4 c = *reinterpret_cast<const float4 *>(Cin + off);
        v.x *= c.x; v.y *= c.y; v.z *= c.z; v.w *= c.w;
        *reinterpret_cast<float4 *>(outp + off) = v;
      } else {
        *reinterpret_cast<float4 *>(outp + off) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ output layer
// out[m] = W_out(2 x d) h[m] + b + offset   (model.py:55, :180-183); one warp per row
__global__ void __launch_bounds__(256) out_layer_fwd_kernel(const float *__restrict__ H, int64_t M, int d,
                                                            const float *__restrict__ W, const float *__restrict__ b,
                                                            float off0, float off1, float2 *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float s0 = 0.f, s1 = 0.f;
  for (int k = lane * 4; k < d; k += 128) {
    const float4 h = *reinterpret_cast<const float4 *>(H + m * d + k);
    const float4 w0 = *reinterpret_cast<const float4 *>(W + k);
    const float4 w1 = *reinterpret_cast<const float4 *>(W + d + k);
    s0 += h.x * w0.x + h.y * w0.y + h.z * w0.z + h.w * w0.w;
    s1 += h.x * w1.x + h.y * w1.y + h.z * w1.z + h.w * w1.w;
  }
  s0 = warp_sum_f(s0); s1 = warp_sum_f(s1);
  if (lane == 0) out[m] = make_float2(fadd(fadd(s0, b[0]), off0), fadd(fadd(s1, b[1]), off1));
}

// dpre_last[m,n] = (g[m,0] W[0,n] + g[m,1] W[1,n]) * cos_last[m,n]
__global__ void __launch_bounds__(256) out_layer_bwd_kernel(const float2 *__restrict__ g, int64_t M, int d,
                                                            const float *__restrict__ W, const float *__restrict__ Cs,
                                                            float *__restrict__ D) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of D per thread
  const int64_t total = M * (d >> 2);
  if (idx >= total) return;
  const int64_t m = idx / (d >> 2);
  const int n = (int)(idx - m * (d >> 2)) * 4;
  const float2 gg = g[m];
  const float4 w0 = *reinterpret_cast<const float4 *>(W + n), w1 = *reinterpret_cast<const float4 *>(W + d + n);
  const float4 c = *reinterpret_cast<const float4 *>(Cs + m * d + n);
  float4 o;
  o.x = (gg.x * w0.x + gg.y * w1.x) * c.x; o.y = (gg.x * w0.y + gg.y * w1.y) * c.y;
  o.z = (gg.x * w0.z + gg.y * w1.z) * c.z; o.w = (gg.x * w0.w + gg.y * w1.w) * c.w;
  *reinterpret_cast<float4 *>(D + m * d + n) = o;
}

// partial[chunk][0..1][n] = sum_{m in chunk} g[m,o] h[m,n];  partial_b[chunk][0..1] = sum g[m,o]
__global__ void __launch_bounds__(128) out_wgrad_partial_kernel(const float2 *__restrict__ g,
                                                                const float *__restrict__ H, int64_t M, int d,
                                                                int64_t rows_per_chunk, float *__restrict__ part,
                                                                float *__restrict__ part_b) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = (r0 + rows_per_chunk < M) ? r0 + rows_per_chunk : M;
  float sg0 = 0.f, sg1 = 0.f;
  for (int n = threadIdx.x * 4; n < d; n += blockDim.x * 4) {
    float4 a0 = make_float4(0, 0, 0, 0), a1 = make_float4(0, 0, 0, 0);
    for (int64_t m = r0; m < r1; ++m) {
      const float2 gg = g[m];
      const float4 h = *reinterpret_cast<const float4 *>(H + m * d + n);
      a0.x += gg.x * h.x; a0.y += gg.x * h.y; a0.z += gg.x * h.z; a0.w += gg.x * h.w;
      a1.x += gg.y * h.x; a1.y += gg.y * h.y; a1.z += gg.y * h.z; a1.w += gg.y * h.w;
      if (n == 0) { sg0 += gg.x; sg1 += gg.y; }
    }
    float *p = part + (int64_t)blockIdx.x * 2 * d;
    *reinterpret_cast<float4 *>(p + n) = a0;
    *reinterpret_cast<float4 *>(p + d + n) = a1;
  }
  if (threadIdx.x == 0) { part_b[blockIdx.x * 2] = sg0; part_b[blockIdx.x * 2 + 1] = sg1; }
}

// partial[chunk][n] = sum_{m in chunk} D[m,n]   (bias gradients of the hidden layers)
__global__ void __launch_bounds__(128) colsum_partial_kernel(const float *__restrict__ D, int64_t M, int d,
                                                             int64_t rows_per_chunk, float *__restrict__ part) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = (r0 + rows_per_chunk < M) ? r0 + rows_per_chunk : M;
  for (int n = threadIdx.x * 4; n < d; n += blockDim.x * 4) {
    float4 a = make_float4(0, 0, 0, 0);
    for (int64_t m = r0; m < r1; ++m) {
      const float4 v = *reinterpret_cast<const float4 *>(D + m * d + n);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4 *>(part + (int64_t)blockIdx.x * d + n) = a;
  }
}

// out[i] = sum_c part[c*count + i]  (fixed order: deterministic)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float *__restrict__ part, int nparts,
                                                              int64_t count, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int c = 0; c < nparts; ++c) s += part[(int64_t)c * count + i];
  out[i] = s;
}

// ------------------------------------------------------------------------------------------ workspace layout
struct F32Ws {
  float *enc;        // [M,84]
  float *H[16];      // activations per hidden layer (train) or 2 ping-pong buffers (inference)
  float *C[16];      // cosines per hidden layer (train only)
  float *D[2];       // backward ping-pong
  float *part;       // split-K / column-sum partials
  int64_t bytes;
};
constexpr int kMaxSplit = 64;
constexpr int kColChunks = 256;

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

inline F32Ws f32_layout(void *base, int64_t M, int n_hidden, int d, int train) {
  F32Ws w{};
  char *p = reinterpret_cast<char *>(base);
  int64_t off = 0;
  auto take = [&](int64_t nfloat) { float *r = reinterpret_cast<float *>(p + off); off += align_up(nfloat * 4, 256); return r; };
  w.enc = take(M * ENC_OUT);
  const int64_t act = M * d;
  if (train) {
    for (int l = 0; l < n_hidden; ++l) w.H[l] = take(act);
    for (int l = 0; l < n_hidden; ++l) w.C[l] = take(act);
    w.D[0] = take(act); w.D[1] = take(act);
    const int64_t wmax = (int64_t)d * (d > ENC_OUT ? d : ENC_OUT);
    const int64_t pa = (int64_t)kMaxSplit * wmax, pb = (int64_t)kColChunks * 2 * d + kColChunks * 2;
    w.part = take(pa > pb ? pa : pb);
  } else {
    w.H[0] = take(act); w.H[1] = take(act);
  }
  w.bytes = off;
  return w;
}

}  // namespace snf

using namespace snf;

extern "C" int64_t snf_mlp_pack_bytes(void);
int64_t snf_mlp_bf16_ws_bytes(int64_t M, int train, int x3);   // snf_mlp_bf16.cu

extern "C" int64_t snf_mlp_ws_bytes(int64_t M, int n_hidden, int d_filter, int mode, int train) {
  if (M < 0 || n_hidden <= 0 || n_hidden > 16 || d_filter <= 0) return SNF_E_ARG;
  if (M == 0) return 256;
  if (mode == 1 || mode == 2) return snf_mlp_bf16_ws_bytes(M, train, mode == 2);
  return f32_layout(nullptr, M, n_hidden, d_filter, train).bytes;
}

extern "C" int snf_mlp_fwd_f32(const float *x, int64_t M, const float *const *W, const float *const *B,
                               int n_hidden, int d, float off0, float off1, float *out, void *ws, int train,
                               void *stream) {
  if (M == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(x); SNF_CHECK_PTR(W); SNF_CHECK_PTR(B); SNF_CHECK_PTR(out); SNF_CHECK_PTR(ws);
  SNF_CHECK_ALIGN(x, 16); SNF_CHECK_ALIGN(out, 8); SNF_CHECK_ALIGN(ws, 256);
  if (M < 0 || n_hidden < 1 || n_hidden > 16) return SNF_E_ARG;
  if (d % 4 != 0 || d < 4) return SNF_E_SHAPE;
  for (int l = 0; l <= n_hidden; ++l) { SNF_CHECK_PTR(W[l]); SNF_CHECK_PTR(B[l]); SNF_CHECK_ALIGN(W[l], 16); SNF_CHECK_ALIGN(B[l], 16); }
  cudaStream_t st = (cudaStream_t)stream;
  F32Ws w = f32_layout(ws, M, n_hidden, d, train);
  encode_kernel<<<(unsigned)ceil_div64(M * 44, 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(x), M, w.enc);
  const dim3 grid((unsigned)((d + BN - 1) / BN), (unsigned)ceil_div64(M, BM), 1);
  const float *in = w.enc;
  int kin = ENC_OUT;
  for (int l = 0; l < n_hidden; ++l) {
    float *h = train ? w.H[l] : w.H[l & 1];
    float *c = train ? w.C[l] : nullptr;
    sgemm_kernel<0, 0, 0><<<grid, 256, 0, st>>>(in, kin, W[l], kin, M, d, kin, kin, B[l], nullptr, h, c, d);
    in = h; kin = d;
  }
  out_layer_fwd_kernel<<<(unsigned)ceil_div64(M, 8), 256, 0, st>>>(in, M, d, W[n_hidden], B[n_hidden], off0, off1,
                                                                 reinterpret_cast<float2 *>(out));
  count_launch(2 + n_hidden);
  return launch_status();
}

extern "C" int snf_mlp_bwd_f32(const float *x, int64_t M, const float *const *W, int n_hidden, int d,
                               const float *grad_out, void *ws, float *const *gW, float *const *gB, void *stream) {
  SNF_CHECK_PTR(x); SNF_CHECK_PTR(W); SNF_CHECK_PTR(grad_out); SNF_CHECK_PTR(ws); SNF_CHECK_PTR(gW); SNF_CHECK_PTR(gB);
  SNF_CHECK_ALIGN(grad_out, 8); SNF_CHECK_ALIGN(ws, 256);
  if (M <= 0 || n_hidden < 1 || n_hidden > 16) return SNF_E_ARG;
  if (d % 4 != 0 || d < 4) return SNF_E_SHAPE;
  for (int l = 0; l <= n_hidden; ++l) { SNF_CHECK_PTR(W[l]); SNF_CHECK_PTR(gW[l]); SNF_CHECK_PTR(gB[l]); SNF_CHECK_ALIGN(gW[l], 16); }
  cudaStream_t st = (cudaStream_t)stream;
  F32Ws w = f32_layout(ws, M, n_hidden, d, 1);
  const float2 *g = reinterpret_cast<const float2 *>(grad_out);
  int launches = 0;
  // ---- output layer
  const int64_t rows_per_chunk = ceil_div64(M, kColChunks);
  const int nchunks = (int)ceil_div64(M, rows_per_chunk);
  float *part_b = w.part + (int64_t)kColChunks * 2 * d;
  out_wgrad_partial_kernel<<<nchunks, 128, 0, st>>>(g, w.H[n_hidden - 1], M, d, rows_per_chunk, w.part, part_b);
  reduce_partials_kernel<<<(unsigned)ceil_div64(2 * d, 256), 256, 0, st>>>(w.part, nchunks, 2 * d, gW[n_hidden]);
  reduce_partials_kernel<<<1, 256, 0, st>>>(part_b, nchunks, 2, gB[n_hidden]);
  int cur = 0;
  out_layer_bwd_kernel<<<(unsigned)ceil_div64(M * (d / 4), 256), 256, 0, st>>>(g, M, d, W[n_hidden], w.C[n_hidden - 1], w.D[cur]);
  launches += 4;
  // ---- hidden layers, last to first
  int64_t k_per_split = ceil_div64(M, kMaxSplit);
  k_per_split = (k_per_split + BK - 1) / BK * BK;
  if (k_per_split < 1024) k_per_split = 1024;
  const int splits = (int)ceil_div64(M, k_per_split);
  for (int l = n_hidden - 1; l >= 0; --l) {
    const float *hin = l > 0 ? w.H[l - 1] : w.enc;
    const int kin = l > 0 ? d : ENC_OUT;
    // gW[l][o,i] = sum_p D[p,o] * hin[p,i]
    const dim3 gw((unsigned)((kin + BN - 1) / BN), (unsigned)((d + BM - 1) / BM), (unsigned)splits);
    sgemm_kernel<1, 1, 2><<<gw, 256, 0, st>>>(w.D[cur], d, hin, kin, d, kin, M, k_per_split, nullptr, nullptr, w.part, nullptr, kin);
    reduce_partials_kernel<<<(unsigned)ceil_div64((int64_t)d * kin, 256), 256, 0, st>>>(w.part, splits, (int64_t)d * kin, gW[l]);
    colsum_partial_kernel<<<nchunks, 128, 0, st>>>(w.D[cur], M, d, rows_per_chunk, w.part);
    reduce_partials_kernel<<<(unsigned)ceil_div64(d, 256), 256, 0, st>>>(w.part, nchunks, d, gB[l]);
    launches += 4;
    if (l > 0) {   // D_prev[p,i] = (sum_o D[p,o] W[l][o,i]) * cos_{l-1}[p,i]
      const dim3 gd((unsigned)((d + BN - 1) / BN), (unsigned)ceil_div64(M, BM), 1);
      sgemm_kernel<0, 1, 1><<<gd, 256, 0, st>>>(w.D[cur], d, W[l], d, M, d, d, d, nullptr, w.C[l - 1], w.D[cur ^ 1], nullptr, d);
      cur ^= 1;
      ++launches;
    }
  }
  count_launch(launches);
  return launch_status();
}
