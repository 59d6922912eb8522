// Front-to-back compositing heads, their analytic backward, the render epilogue, SimpleStar and the loss.
// All HBM-bound: one warp per ray, coalesced row loads, warp-shuffle scans (double accumulators, which is
// what torch's CPU cumsum/cumprod use), results staged in the warp's slice of shared memory.
//   K5  composite_emission_{fwd,bwd} : sunerf/rendering/emission.py:14-54, base_tracing.py:135-156
//   K6  composite_dt_{fwd,bwd}       : sunerf/rendering/density_temperature.py:192-271
//   render_epilogue_kernel           : sunerf/rendering/base_tracing.py:91-111, :43-44; density_temperature.py:273-274
//   simple_star_kernel               : sunerf/model/stellar_model.py:53-102
//   loss kernels                     : sunerf/model/sunerf.py:98-131, 173-206; sunerf/train/scaling.py:17-28
#include "snf_common.cuh"
#include <type_traits>

namespace snf {

constexpr int kRayWarps = 4;  // warps (= rays) per CTA for the per-ray kernels

// ------------------------------------------------------------------------------------------------ K5
// One warp per ray, lane <-> sample c*32 + lane of chunk c.  Every global load of the ray is issued before the first
// use (S = 192: 2.3 KB in flight per warp), the whole ray lives in registers, neighbours come from shuffles; the
// only serial dependence is the carry of the exclusive product between the NCH chunk scans.
// dz of a sample: z[j] - z[j-1], the first one z[1] - z[0] (emission.py:24-29).
template <int NCH>
__device__ __forceinline__ void ray_dz(const float (&zr)[NCH], int lane, float dnorm, int S, float (&dz)[NCH]) {
  const float z1 = __shfl_sync(kFull, zr[0], 1);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float up = __shfl_up_sync(kFull, zr[c], 1);
    const float prev31 = __shfl_sync(kFull, zr[c > 0 ? c - 1 : 0], 31);
    const float d = lane > 0 ? fsub(zr[c], up) : (c > 0 ? fsub(zr[c], prev31) : fsub(z1, zr[0]));
    dz[c] = fmul(d, dnorm);
  }
}

template <int NCH>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_emission_fwd_kernel(const float2 *__restrict__ raw, const float *__restrict__ z,
                                  const float *__restrict__ rays_d, int64_t N, int S, float *__restrict__ image,
                                  float *__restrict__ weights, float *__restrict__ absorption) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp;
  if (ray >= N) return;
  float zr[NCH];
  float2 rw[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = c * 32 + lane;
    zr[c] = j < S ? __ldcs(z + ray * S + j) : 0.f;
    rw[c] = j < S ? __ldcs(raw + ray * S + j) : make_float2(0.f, 0.f);
  }
  const float d0 = rays_d[3 * ray], d1 = rays_d[3 * ray + 1], d2 = rays_d[3 * ray + 2];
  const float dnorm = __fsqrt_rn(sum3(fmul(d0, d0), fmul(d1, d1), fmul(d2, d2)));      // :29
  float dz[NCH], P[NCH];
  ray_dz<NCH>(zr, lane, dnorm, S, dz);
  double carry = 1.0, isum = 0.0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = c * 32 + lane;
    const bool ok = j < S;
    float E = 0.f, f = 1.f;
    if (ok) {
      E = fmul(expf(rw[c].x), dz[c]);                                                    // :34
      const float a = expf(fmul(-fmaxf(rw[c].y, 0.f), dz[c]));                           // :37
      __stcs(absorption + ray * S + j, a);
      f = fadd(a, 1e-10f);                                                               // :43
    }
    const double incl = warp_incl_prod((double)f, lane);
    double excl = shfl_up_d(incl, 1);
    if (lane == 0) excl = 1.0;
    const float T = (float)(carry * excl);     // exclusive cumprod, double accumulate -> float per prefix
    P[c] = fmul(E, T);                                                                   // :46
    isum += (double)P[c];
    carry *= __shfl_sync(kFull, incl, 31);
  }
  const float I = (float)warp_sum(isum);                                                 // :48
  if (lane == 0) image[ray] = I;
  const float den = fadd(I, 1e-10f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = c * 32 + lane;
    if (j < S) __stcs(weights + ray * S + j, fdiv(P[c], den));                           // :51-52
  }
}

// dI/draw0[k] = P_k ; dI/draw1[j] = -1[raw1>0] dz_j a_j (sum_{k>j} P_k)/(a_j+1e-10) ; plus dL/da from g_absorption
template <int NCH>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_emission_bwd_kernel(const float2 *__restrict__ raw, const float *__restrict__ z,
                                  const float *__restrict__ rays_d, int64_t N, int S,
                                  const float *__restrict__ g_image, const float *__restrict__ g_abs,
                                  float2 *__restrict__ g_raw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp;
  if (ray >= N) return;
  float zr[NCH], ga_in[NCH];
  float2 rw[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = c * 32 + lane;
    zr[c] = j < S ? __ldcs(z + ray * S + j) : 0.f;
    rw[c] = j < S ? __ldcs(raw + ray * S + j) : make_float2(0.f, 0.f);
    ga_in[c] = (g_abs != nullptr && j < S) ? __ldcs(g_abs + ray * S + j) : 0.f;
  }
  const float d0 = rays_d[3 * ray], d1 = rays_d[3 * ray + 1], d2 = rays_d[3 * ray + 2];
  const float dnorm = __fsqrt_rn(sum3(fmul(d0, d0), fmul(d1, d1), fmul(d2, d2)));
  const float g = g_image[ray];
  float dz[NCH], P[NCH], A[NCH];
  ray_dz<NCH>(zr, lane, dnorm, S, dz);
  double carry = 1.0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const bool ok = c * 32 + lane < S;
    float E = 0.f, f = 1.f, a = 1.f;
    if (ok) {
      E = fmul(expf(rw[c].x), dz[c]);
      a = expf(fmul(-fmaxf(rw[c].y, 0.f), dz[c]));
      f = fadd(a, 1e-10f);
    }
    const double incl = warp_incl_prod((double)f, lane);
    double excl = shfl_up_d(incl, 1);
    if (lane == 0) excl = 1.0;
    const float T = (float)(carry * excl);
    P[c] = ok ? fmul(E, T) : 0.f;
    A[c] = a;
    carry *= __shfl_sync(kFull, incl, 31);
  }
  // reverse pass: exclusive suffix sums of P (fp32 is ample for the 1e-3 gradient tolerance: autograd itself is fp32)
  float rcarry = 0.f;
#pragma unroll
  for (int c = NCH - 1; c >= 0; --c) {
    const int j = c * 32 + lane;
    float suf = P[c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float n = __shfl_down_sync(kFull, suf, d);
      if (lane + d < 32) suf += n;
    }
    suf += rcarry;                                          // inclusive
    const float suf_excl = suf - P[c];
    if (j < S) {
      const float a = A[c];
      const float ga = g * suf_excl / fadd(a, 1e-10f) + ga_in[c];
      float2 o;
      o.x = g * P[c];
      o.y = (rw[c].y > 0.f) ? -ga * dz[c] * a : 0.f;
      __stcs(g_raw + ray * S + j, o);
    }
    rcarry = __shfl_sync(kFull, suf, 0);
  }
}

// ------------------------------------------------------------------------------------------------ K5, S = 64 NC
// Pair layout for the configs' sample counts (64 coarse, 192 = 64 + 128 fine): a chunk is 64 consecutive samples, lane l
// owns samples 64 c + 2 l and 64 c + 2 l + 1.  Against the one-sample-per-lane kernels above that is half the
// double-precision scans (one per 64 samples: the pair's product / sum is formed in registers first), half the neighbour
// shuffles, no bounds predicates, and 8 / 16-byte loads and stores that a warp still issues fully coalesced.  The generic
// kernels stay for every other S.  Same arithmetic per sample; the exclusive product is T_j = float(prod_{k<j} f_k) in
// double with a different association (1e-16 before the rounding to float, as in the generic kernel's tree scan).
template <int NC>
struct PairRay {
  float dz[2 * NC], E[2 * NC], a[2 * NC], P[2 * NC];
  float rawy[2 * NC];
};

// loads of one ray (all issued before the first use), dz (emission.py:24-29), E, a and P = E T with the exclusive product T
template <int NC>
__device__ __forceinline__ void pair_ray_forward(const float2 *__restrict__ raw, const float *__restrict__ z, int64_t ray,
                                                 int lane, float dnorm, PairRay<NC> &r) {
  constexpr int S = 64 * NC;
  float2 zz[NC];
  float4 rw[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    zz[c] = __ldcs(reinterpret_cast<const float2 *>(z + ray * S) + c * 32 + lane);
    rw[c] = __ldcs(reinterpret_cast<const float4 *>(raw + ray * S) + c * 32 + lane);
  }
  double carry = 1.0;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float up = __shfl_up_sync(kFull, zz[c].y, 1);
    const float prev31 = __shfl_sync(kFull, zz[c > 0 ? c - 1 : 0].y, 31);
    const float d0 = lane > 0 ? fsub(zz[c].x, up) : (c > 0 ? fsub(zz[c].x, prev31) : fsub(zz[0].y, zz[0].x));
    const float dz0 = fmul(d0, dnorm), dz1 = fmul(fsub(zz[c].y, zz[c].x), dnorm);
    const float E0 = fmul(expf(rw[c].x), dz0), E1 = fmul(expf(rw[c].z), dz1);              // emission.py:34
    const float a0 = expf(fmul(-fmaxf(rw[c].y, 0.f), dz0)), a1 = expf(fmul(-fmaxf(rw[c].w, 0.f), dz1));   // :37
    const double f0 = (double)fadd(a0, 1e-10f), f1 = (double)fadd(a1, 1e-10f);             // :43
    const double incl = warp_incl_prod(f0 * f1, lane);
    double excl = shfl_up_d(incl, 1);
    if (lane == 0) excl = 1.0;
    const double t0 = carry * excl;
    r.dz[2 * c] = dz0; r.dz[2 * c + 1] = dz1;
    r.E[2 * c] = E0; r.E[2 * c + 1] = E1;
    r.a[2 * c] = a0; r.a[2 * c + 1] = a1;
    r.rawy[2 * c] = rw[c].y; r.rawy[2 * c + 1] = rw[c].w;
    r.P[2 * c] = fmul(E0, (float)t0);                                                      // :46
    r.P[2 * c + 1] = fmul(E1, (float)(t0 * f0));
    carry *= __shfl_sync(kFull, incl, 31);
  }
}

template <int NC>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_emission_fwd_pair_kernel(const float2 *__restrict__ raw, const float *__restrict__ z,
                                       const float *__restrict__ rays_d, int64_t N, float *__restrict__ image,
                                       float *__restrict__ weights, float *__restrict__ absorption) {
  constexpr int S = 64 * NC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp;
  if (ray >= N) return;
  const float d0 = rays_d[3 * ray], d1 = rays_d[3 * ray + 1], d2 = rays_d[3 * ray + 2];
  const float dnorm = __fsqrt_rn(sum3(fmul(d0, d0), fmul(d1, d1), fmul(d2, d2)));          // :29
  PairRay<NC> r;
  pair_ray_forward<NC>(raw, z, ray, lane, dnorm, r);
  double isum = 0.0;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    __stcs(reinterpret_cast<float2 *>(absorption + ray * S) + c * 32 + lane, make_float2(r.a[2 * c], r.a[2 * c + 1]));
    isum += (double)r.P[2 * c] + (double)r.P[2 * c + 1];
  }
  const float I = (float)warp_sum(isum);                                                   // :48
  if (lane == 0) image[ray] = I;
  const float den = fadd(I, 1e-10f);
#pragma unroll
  for (int c = 0; c < NC; ++c)
    __stcs(reinterpret_cast<float2 *>(weights + ray * S) + c * 32 + lane,
           make_float2(fdiv(r.P[2 * c], den), fdiv(r.P[2 * c + 1], den)));                 // :51-52
}

template <int NC>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_emission_bwd_pair_kernel(const float2 *__restrict__ raw, const float *__restrict__ z,
                                       const float *__restrict__ rays_d, int64_t N, const float *__restrict__ g_image,
                                       const float *__restrict__ g_abs, float2 *__restrict__ g_raw) {
  constexpr int S = 64 * NC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp;
  if (ray >= N) return;
  float2 ga_in[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    ga_in[c] = g_abs != nullptr ? __ldcs(reinterpret_cast<const float2 *>(g_abs + ray * S) + c * 32 + lane) : make_float2(0.f, 0.f);
  const float d0 = rays_d[3 * ray], d1 = rays_d[3 * ray + 1], d2 = rays_d[3 * ray + 2];
  const float dnorm = __fsqrt_rn(sum3(fmul(d0, d0), fmul(d1, d1), fmul(d2, d2)));
  const float g = g_image[ray];
  PairRay<NC> r;
  pair_ray_forward<NC>(raw, z, ray, lane, dnorm, r);
  // reverse pass: exclusive suffix sums of P (fp32, as in the generic kernel)
  float rcarry = 0.f;
#pragma unroll
  for (int c = NC - 1; c >= 0; --c) {
    const float P0 = r.P[2 * c], P1 = r.P[2 * c + 1];
    float suf = P0 + P1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float n = __shfl_down_sync(kFull, suf, d);
      if (lane + d < 32) suf += n;
    }
    suf += rcarry;                                   // inclusive of this lane's pair
    const float ex1 = suf - (P0 + P1), ex0 = ex1 + P1;   // sum_{k > j} P_k for the pair's second / first sample
    const float a0 = r.a[2 * c], a1 = r.a[2 * c + 1];
    const float ga0 = g * ex0 / fadd(a0, 1e-10f) + ga_in[c].x, ga1 = g * ex1 / fadd(a1, 1e-10f) + ga_in[c].y;
    float4 o;
    o.x = g * P0;
    o.y = r.rawy[2 * c] > 0.f ? -ga0 * r.dz[2 * c] * a0 : 0.f;
    o.z = g * P1;
    o.w = r.rawy[2 * c + 1] > 0.f ? -ga1 * r.dz[2 * c + 1] * a1 : 0.f;
    __stcs(reinterpret_cast<float4 *>(g_raw + ray * S) + c * 32 + lane, o);
    rcarry = __shfl_sync(kFull, suf, 0);
  }
}

// ------------------------------------------------------------------------------------------------ K6
struct DtTables {
  float x[SNF_TABLE_LEN];
  float y[SNF_N_AIA][SNF_TABLE_LEN];
  float slope[SNF_N_AIA][SNF_TABLE_LEN - 1];
  float2 ys[SNF_N_AIA][SNF_TABLE_LEN - 1];   // (y, slope) of a segment side by side: one 8-byte load per lookup
  float kappa[SNF_N_AIA];      // relu(log_abs)
  float kappa_on[SNF_N_AIA];   // 1[log_abs > 0]
};

__device__ __forceinline__ void dt_load_tables(DtTables *t, const float *table_x, const float *table_y,
                                               const float *log_abs) {
  for (int i = threadIdx.x; i < SNF_TABLE_LEN; i += blockDim.x) t->x[i] = table_x[i];
  for (int i = threadIdx.x; i < SNF_N_AIA * SNF_TABLE_LEN; i += blockDim.x)
    t->y[i / SNF_TABLE_LEN][i % SNF_TABLE_LEN] = table_y[i];
  for (int i = threadIdx.x; i < SNF_N_AIA * (SNF_TABLE_LEN - 1); i += blockDim.x) {
    const int k = i / (SNF_TABLE_LEN - 1), s = i % (SNF_TABLE_LEN - 1);
    // xitorch LinearInterp1D: per-segment slope (y[1:]-y[:-1])/(x[1:]-x[:-1])
    const float sl = fdiv(fsub(table_y[k * SNF_TABLE_LEN + s + 1], table_y[k * SNF_TABLE_LEN + s]),
                          fsub(table_x[s + 1], table_x[s]));
    t->slope[k][s] = sl;
    t->ys[k][s] = make_float2(table_y[k * SNF_TABLE_LEN + s], sl);
  }
  if (threadIdx.x < SNF_N_AIA) {
    const float la = log_abs[threadIdx.x];
    t->kappa[threadIdx.x] = fmaxf(la, 0.f);                         // density_temperature.py:256
    t->kappa_on[threadIdx.x] = la > 0.f ? 1.f : 0.f;
  }
  __syncthreads();
}

__device__ __forceinline__ int dt_channel(float wl) {   // wavelength value -> row of the response table
  const int w = (int)wl;
  switch (w) {
    case 94: return 0; case 131: return 1; case 171: return 2; case 193: return 3;
    case 211: return 4; case 304: return 5; case 335: return 6; default: return -1;
  }
}

// the same mapping without branches (lane c decodes channel c: the switch above would serialise seven paths)
__device__ __forceinline__ int dt_channel_sel(float wl) {
  const int w = (int)wl;
  int k = -1;
  k = w == 94 ? 0 : k; k = w == 131 ? 1 : k; k = w == 171 ? 2 : k; k = w == 193 ? 3 : k;
  k = w == 211 ? 4 : k; k = w == 304 ? 5 : k; k = w == 335 ? 6 : k;
  return k;
}

// segment lookup shared by all channels of a sample: idxl in [0,99] or -1 when theta is outside [x0, x100]
__device__ __forceinline__ int dt_segment(const DtTables *t, float th) {
  if (!(th >= t->x[0]) || !(th <= t->x[SNF_TABLE_LEN - 1])) return -1;   // extrap=0 (also NaN)
  int g = (int)((th - t->x[0]) * 20.f);                 // grid step 0.05; corrected against the table below
  g = min(max(g, 0), SNF_TABLE_LEN - 2);
  // searchsorted(left): idxr = #{x < th} clamped to [1,100]; idxl = idxr-1, i.e. x[idxl] < th <= x[idxl+1]
  while (g > 0 && t->x[g] >= th) --g;
  while (g < SNF_TABLE_LEN - 2 && t->x[g + 1] < th) ++g;
  return g;
}

// K6 layout: persistent CTAs (the response tables and their slopes are built once per CTA, not once per 4 rays),
// one warp per ray, the ray register-resident as in K5.  The optical depth of channel c is kappa_c * B / 2 with
//   B[k] = cumsum_k (z[k+1]-z[k]) (rho[k] + rho[k+1])        (cumulative_trapezoid of rho, :261)
// so ONE double-precision scan per ray serves all C channels (the reference multiplies by kappa_c inside the sum;
// factoring it out moves the result by ~1e-7 relative, the gate is 1e-5).
template <int NCH>
struct DtRay {
  float z[NCH], rho[NCH], dxq[NCH], dzn[NCH], B[NCH];   // dzn[k] = z[k+1]-z[k]; B = 2 x cumulative trapezoid of rho
  int seg[NCH];
};

// value of the next sample (lane + 1), crossing into the next chunk's lane 0
template <int NCH>
__device__ __forceinline__ void next_sample(const float (&v)[NCH], int lane, float (&nx)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float dn = __shfl_down_sync(kFull, v[c], 1);
    const float first_next = __shfl_sync(kFull, v[c + 1 < NCH ? c + 1 : c], 0);
    nx[c] = lane < 31 ? dn : first_next;
  }
}
template <int NCH>
__device__ __forceinline__ void prev_sample(const float (&v)[NCH], int lane, float (&pv)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float up = __shfl_up_sync(kFull, v[c], 1);
    const float last_prev = __shfl_sync(kFull, v[c > 0 ? c - 1 : c], 31);
    pv[c] = lane > 0 ? up : last_prev;
  }
}

// 2^x on the MUFU: the attenuation exp(-A) of every (channel, sample) pair.  Relative error 2^-22 + |x| 2^-24 - at the
// optical depths that still contribute (exp(-A) > 1e-5 of a pixel, A < 12) below 1.5e-6 of a term, against a 1e-5 gate on
// the pixel; the accurate expf it replaces was a third of the kernel's instructions (7 channels x S calls per ray).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NCH>
__device__ __forceinline__ void dt_ray_setup(DtRay<NCH> &r, const DtTables *tab, const float (&zr)[NCH],
                                             const float2 (&v)[NCH], int lane, int S) {
  float rn[NCH], zn[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    r.z[c] = zr[c];
    r.rho[c] = expf(fmaxf(v[c].x, 0.f));                                      // :237
    const float th = fmaxf(v[c].y, 0.f);                                      // :241
    const int sg = (c * 32 + lane < S) ? dt_segment(tab, th) : -1;
    r.seg[c] = sg;
    r.dxq[c] = sg >= 0 ? fsub(th, tab->x[sg]) : 0.f;
  }
  next_sample<NCH>(r.rho, lane, rn);
  next_sample<NCH>(r.z, lane, zn);
  double carry = 0.0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const bool ok = c * 32 + lane < S - 1;
    r.dzn[c] = ok ? fsub(zn[c], r.z[c]) : 0.f;
    const double term = ok ? (double)fmul(r.dzn[c], fadd(r.rho[c], rn[c])) : 0.0;
    const double inc = warp_incl_sum(term, lane) + carry;
    r.B[c] = (float)inc;
    carry = __shfl_sync(kFull, inc, 31);
  }
}

template <int NCH>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_dt_fwd_kernel(const float2 *__restrict__ inf, const float *__restrict__ z,
                            const float *__restrict__ wavelengths, int64_t N, int S, int C,
                            const float *__restrict__ log_abs, const float *__restrict__ vol_c,
                            const float *__restrict__ table_x, const float *__restrict__ table_y, float F,
                            float *__restrict__ image, float *__restrict__ weights, float *__restrict__ regq) {
  __shared__ DtTables tabs;
  const DtTables *tab = &tabs;
  dt_load_tables(&tabs, table_x, table_y, log_abs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float vc = vol_c[0];
  for (int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp; ray < N; ray += (int64_t)gridDim.x * kRayWarps) {
    float zr[NCH];
    float2 v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = c * 32 + lane;
      zr[c] = j < S ? __ldcs(z + ray * S + j) : 0.f;
      v[c] = j < S ? __ldcs(inf + ray * S + j) : make_float2(0.f, 0.f);
    }
    double qsum = 0.0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = c * 32 + lane;
      const float q = fmaxf(v[c].x, 0.f);
      if (j < S) { __stcs(regq + ray * S + j, q); qsum += (double)q; }         // :271
    }
    const float den = fadd((float)warp_sum(qsum), 1e-10f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = c * 32 + lane;
      if (j < S) __stcs(weights + ray * S + j, fdiv(fmaxf(v[c].x, 0.f), den));  // :268-269
    }
    DtRay<NCH> r;
    dt_ray_setup<NCH>(r, tab, zr, v, lane, S);
    // The trapezoid over z[0..S-2] (:265) as a weighted sum of its nodes: J = sum_k w_k tau_k with
    // w_k = (dz_{k-1} [k >= 1] + dz_k [k <= S-3]) / 2 - the same sum as sum(dz (left + right)) / 2 without the
    // neighbour exchange per channel; rho^2 and B / 2 are formed once per sample, not once per channel.
    float dzp[NCH], wk[NCH], rho2[NCH], Bh[NCH];
    prev_sample<NCH>(r.dzn, lane, dzp);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = ch * 32 + lane;
      wk[ch] = j <= S - 2 ? 0.5f * ((j >= 1 ? dzp[ch] : 0.f) + (j <= S - 3 ? r.dzn[ch] : 0.f)) : 0.f;
      rho2[ch] = fmul(r.rho[ch], r.rho[ch]);                                    // :263
      Bh[ch] = 0.5f * r.B[ch];                                                  // :261 (exact halving)
    }
    for (int c = 0; c < C; ++c) {
      const int k = dt_channel(wavelengths[ray * C + c]);
      if (k < 0) {   // channel absent: response and absorption stay 0 (:243, :251) -> image 0 * vol_c * F
        if (lane == 0) image[ray * C + c] = fmul(fmul(0.f, vc), F);
        continue;
      }
      const float nk = -tab->kappa[k] * 1.4426950408889634f;                    // exp(-kappa B/2) = 2^(nk B/2)
      float part = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int sg = r.seg[ch];
        const float2 ys = tab->ys[k][sg >= 0 ? sg : 0];
        const float R = sg >= 0 ? fmaf(r.dxq[ch], ys.y, ys.x) : 0.f;            // :248 linear interpolation, 0 outside
        part = fmaf(wk[ch], ex2_approx(nk * Bh[ch]) * (rho2[ch] * R), part);    // :264-265
      }
      const float J = warp_sum_f(part);
      if (lane == 0) image[ray * C + c] = fmul(fmul(J, vc), F);
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_dt_bwd_kernel(const float2 *__restrict__ inf, const float *__restrict__ z,
                            const float *__restrict__ wavelengths, int64_t N, int S, int C,
                            const float *__restrict__ log_abs, const float *__restrict__ vol_c,
                            const float *__restrict__ table_x, const float *__restrict__ table_y, float F,
                            const float *__restrict__ g_image, const float *__restrict__ g_regq,
                            float2 *__restrict__ g_inf, float *__restrict__ g_log_abs, float *__restrict__ g_vol_c) {
  __shared__ DtTables tabs;
  __shared__ float blk_acc[8];   // 7 kappa grads + vol_c grad of this CTA's rays
  const DtTables *tab = &tabs;
  if (threadIdx.x < 8) blk_acc[threadIdx.x] = 0.f;
  dt_load_tables(&tabs, table_x, table_y, log_abs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float vc = vol_c[0];
  float acc_k[SNF_N_AIA], gvc = 0.f;
#pragma unroll
  for (int k = 0; k < SNF_N_AIA; ++k) acc_k[k] = 0.f;
  for (int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp; ray < N; ray += (int64_t)gridDim.x * kRayWarps) {
    float zr[NCH], greg[NCH];
    float2 v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = c * 32 + lane;
      zr[c] = j < S ? __ldcs(z + ray * S + j) : 0.f;
      v[c] = j < S ? __ldcs(inf + ray * S + j) : make_float2(0.f, 0.f);
      greg[c] = (g_regq != nullptr && j < S) ? __ldcs(g_regq + ray * S + j) : 0.f;
    }
    DtRay<NCH> r;
    dt_ray_setup<NCH>(r, tab, zr, v, lane, S);
    // trapezoid node weights over z[0..S-2]: w_k = (dz_{k-1} [k>=1] + dz_k [k<=S-3]) / 2  (J = sum_k w_k tau_k, as in the forward)
    float dzp[NCH], wk[NCH], drho[NCH], dth[NCH], dB[NCH], rho2[NCH], Bh[NCH];
    prev_sample<NCH>(r.dzn, lane, dzp);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = ch * 32 + lane;
      wk[ch] = j <= S - 2 ? 0.5f * ((j >= 1 ? dzp[ch] : 0.f) + (j <= S - 3 ? r.dzn[ch] : 0.f)) : 0.f;
      rho2[ch] = fmul(r.rho[ch], r.rho[ch]);
      Bh[ch] = 0.5f * r.B[ch];
      drho[ch] = dth[ch] = dB[ch] = 0.f;
    }
    for (int c = 0; c < C; ++c) {
      const int k = dt_channel(wavelengths[ray * C + c]);
      if (k < 0) continue;   // image is the constant 0: no gradient to anything but vol_c (0 * F)
      const float kap = tab->kappa[k];
      const float nk = -kap * 1.4426950408889634f;
      const float gi = g_image[ray * C + c];
      const float Gc = gi * vc * F;            // dL/dJ with I = J * vol_c * F
      float part = 0.f, dk = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int sg = r.seg[ch];
        const float2 ys = tab->ys[k][sg >= 0 ? sg : 0];
        const float R = sg >= 0 ? fmaf(r.dxq[ch], ys.y, ys.x) : 0.f;
        const float eA = ex2_approx(nk * Bh[ch]);
        const float gw = Gc * wk[ch];          // 0 beyond the last trapezoid node
        const float tau = eA * (rho2[ch] * R);
        part = fmaf(wk[ch], tau, part);
        // dL/dA_k = -Gc w_k tau_k with A = kappa B / 2: dL/dB_k += kappa/2 dL/dA_k ; dL/dkappa += B_k/2 dL/dA_k
        const float dA = -gw * tau;
        dB[ch] = fmaf(0.5f * kap, dA, dB[ch]);
        dk = fmaf(Bh[ch], dA, dk);
        const float dem = gw * eA;             // dL/d em_j
        if (sg >= 0) {
          drho[ch] = fmaf(dem * 2.f * r.rho[ch], R, drho[ch]);
          dth[ch] = fmaf(dem * rho2[ch], ys.y, dth[ch]);
        }
      }
      gvc += gi * warp_sum_f(part) * F;
      dk *= tab->kappa_on[k];
#pragma unroll
      for (int kk = 0; kk < SNF_N_AIA; ++kk) acc_k[kk] += (kk == k) ? dk : 0.f;   // static indices: stays in registers
    }
    // G_i = sum_{k>=i} dL/dB_k = dL/d term_i, term_i = dz_i (rho_i + rho_{i+1}): one suffix scan for all channels
    float G[NCH], Gp[NCH];
    float rcarry = 0.f;
#pragma unroll
    for (int ch = NCH - 1; ch >= 0; --ch) {
      float suf = dB[ch];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float n = __shfl_down_sync(kFull, suf, d);
        if (lane + d < 32) suf += n;
      }
      suf += rcarry;
      G[ch] = suf;
      rcarry = __shfl_sync(kFull, suf, 0);
    }
    prev_sample<NCH>(G, lane, Gp);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = ch * 32 + lane;
      if (j < S) {
        float dr = drho[ch];
        if (j <= S - 2) dr += r.dzn[ch] * G[ch];
        if (j >= 1) dr += dzp[ch] * Gp[ch];
        float2 o;
        o.x = v[ch].x > 0.f ? dr * r.rho[ch] + greg[ch] : 0.f;
        o.y = v[ch].y > 0.f ? dth[ch] : 0.f;
        __stcs(g_inf + ray * S + j, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < SNF_N_AIA; ++k) {
    const float t = warp_sum_f(acc_k[k]);
    if (lane == 0 && t != 0.f) atomicAdd(&blk_acc[k], t);
  }
  if (lane == 0 && gvc != 0.f) atomicAdd(&blk_acc[7], gvc);
  __syncthreads();
  if (threadIdx.x < 7 && blk_acc[threadIdx.x] != 0.f) atomicAdd(&g_log_abs[threadIdx.x], blk_acc[threadIdx.x]);
  if (threadIdx.x == 7 && blk_acc[7] != 0.f) atomicAdd(g_vol_c, blk_acc[7]);
}

// ------------------------------------------------------------------------------------------------ K6, S = 64 NC
// The fast path of the density-temperature head for the configs' sample counts.  The kernel is bound by instruction issue
// (ncu: 2100 warp instructions per ray in the first version of this path, the ready warps "not selected"), so it is about
// instructions:
//   * per ray, not per channel: lane c loads and decodes wavelength c (and dL/dI_c) with the ray, the channel loop gets its
//     table row and kappa by shuffle - no dependent global load and no 25-instruction decode per channel;
//   * everything that does not depend on the channel is folded into per-sample factors once per ray,
//       wr = w_k rho^2 (0 where the temperature is outside the table),  J_c = sum_k wr_k 2^(nk_c B_k / 2) R_c(theta_k),
//     and the table is read through 32-bit shared addresses formed once per sample: the forward inner loop is table load,
//     FMA, product, MUFU (ex2.approx.ftz: one instruction), product, FMA;
//   * the C pixel sums of a ray are reduced together (values handed down a halving tree: 9 shuffles for 8 channels
//     instead of 40), the kappa gradients likewise, straight into the CTA's shared accumulators;
//   * the segment of a temperature on the (verified uniform) log T grid is the arithmetic guess corrected by one
//     branch-free step; any other grid takes the search loops of the generic kernel.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// true when table_x is the 0.05-step grid the arithmetic segment guess assumes (within 2e-3: the guess is then off by at
// most one segment); every thread of the CTA must call it
__device__ __forceinline__ bool dt_grid_is_uniform(const DtTables *t) {
  bool ok = true;
  for (int i = threadIdx.x; i < SNF_TABLE_LEN; i += blockDim.x) ok &= fabsf(t->x[i] - (t->x[0] + 0.05f * (float)i)) < 2e-3f;
  return __syncthreads_and(ok) != 0;
}

constexpr uint32_t kYsRow = (SNF_TABLE_LEN - 1) * sizeof(float2);   // bytes between the (y, slope) rows of two channels
template <uint32_t IMM> __device__ __forceinline__ float2 lds_f32x2_imm(uint32_t addr) {   // [addr + IMM]: the offset rides in the instruction
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(IMM));
  return v;
}

// segment of th: x[g] < th <= x[g+1] (searchsorted left, clamped), -1 outside the table; x_s = shared address of tab->x
__device__ __forceinline__ int dt_segment_fast(const DtTables *tab, uint32_t x_s, float x0, float xN, float th, bool uniform,
                                               float &xg) {
  if (!uniform) {
    const int g = dt_segment(tab, th);
    xg = g >= 0 ? tab->x[g] : 0.f;
    return g;
  }
  int g = (int)((th - x0) * 20.f);
  g = min(max(g, 0), SNF_TABLE_LEN - 2);
  const float2 xa = make_float2(lds_f32(x_s + 4 * g), lds_f32(x_s + 4 * g + 4));
  g -= (xa.x >= th && g > 0) ? 1 : 0;                 // at most one of the two corrections applies
  g += (xa.y < th && g < SNF_TABLE_LEN - 2) ? 1 : 0;
  xg = lds_f32(x_s + 4 * g);
  return (th >= x0 && th <= xN) ? g : -1;             // extrap = 0 (also NaN)
}

// BLOCKED layout of the fast path (S = 32 PER, PER even): lane l owns the PER consecutive samples l PER .. l PER + PER - 1, so
// the whole ray needs ONE double-precision scan (of the lanes' totals; the prefix inside a lane is PER register adds) and one
// neighbour exchange - the pair layout of K5 needed one per 64 samples, and with a single channel the per-ray set-up alone
// held this kernel at 55 % of the HBM peak.  The lane's 4 PER / 8 PER contiguous bytes are read with 8 / 16-byte loads at a
// 4 PER-byte lane stride: three times the L1 wavefronts of a dense access, the same DRAM sectors - the kernel is bound by
// instruction issue, not by L1.
template <int PER>
struct DtBlkRay {
  float rho[PER], dxq[PER], dzn[PER], dzp[PER], Bh[PER], wr[PER];
  uint32_t ysa[PER];   // shared address of the sample's segment in channel 0's (y, slope) row (segment 0 when outside)
};

template <int PER> __device__ __forceinline__ void ld_blk(const float *row, int lane, float (&v)[PER]) {
#pragma unroll
  for (int i = 0; i < PER / 2; ++i) {
    const float2 t = __ldcs(reinterpret_cast<const float2 *>(row) + lane * (PER / 2) + i);
    v[2 * i] = t.x; v[2 * i + 1] = t.y;
  }
}
template <int PER> __device__ __forceinline__ void ld_blk2(const float2 *row, int lane, float2 (&v)[PER]) {
#pragma unroll
  for (int i = 0; i < PER / 2; ++i) {
    const float4 t = __ldcs(reinterpret_cast<const float4 *>(row) + lane * (PER / 2) + i);
    v[2 * i] = make_float2(t.x, t.y); v[2 * i + 1] = make_float2(t.z, t.w);
  }
}
template <int PER> __device__ __forceinline__ void st_blk(float *row, int lane, const float (&v)[PER]) {
#pragma unroll
  for (int i = 0; i < PER / 2; ++i) __stcs(reinterpret_cast<float2 *>(row) + lane * (PER / 2) + i, make_float2(v[2 * i], v[2 * i + 1]));
}

template <int PER>
__device__ __forceinline__ void dt_blk_setup(DtBlkRay<PER> &r, const DtTables *tab, uint32_t x_s, uint32_t ys_s, bool uniform,
                                             const float (&z)[PER], const float2 (&v)[PER], int lane) {
  const float x0 = lds_f32(x_s), xN = lds_f32(x_s + 4 * (SNF_TABLE_LEN - 1));
  bool in[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    r.rho[i] = expf(fmaxf(v[i].x, 0.f));                                      // :237
    const float th = fmaxf(v[i].y, 0.f);                                      // :241
    float xg;
    const int sg = dt_segment_fast(tab, x_s, x0, xN, th, uniform, xg);
    in[i] = sg >= 0;
    r.ysa[i] = ys_s + (uint32_t)(sg >= 0 ? sg : 0) * (uint32_t)sizeof(float2);
    r.dxq[i] = sg >= 0 ? fsub(th, xg) : 0.f;
  }
  // term_k = dz_k (rho_k + rho_{k+1}), B = inclusive cumulative sum (:261, x 2); the last sample of the ray has no successor
  const float zdn = __shfl_down_sync(kFull, z[0], 1), rdn = __shfl_down_sync(kFull, r.rho[0], 1);
  double pre[PER];
  double run = 0.0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const bool last = (i == PER - 1) && lane == 31;
    const float zn = i + 1 < PER ? z[i + 1 < PER ? i + 1 : i] : zdn, rn = i + 1 < PER ? r.rho[i + 1 < PER ? i + 1 : i] : rdn;
    r.dzn[i] = last ? 0.f : fsub(zn, z[i]);
    run += last ? 0.0 : (double)fmul(r.dzn[i], fadd(r.rho[i], rn));
    pre[i] = run;
  }
  const double incl = warp_incl_sum(run, lane);
  double excl = shfl_up_d(incl, 1);
  if (lane == 0) excl = 0.0;
#pragma unroll
  for (int i = 0; i < PER; ++i) r.Bh[i] = 0.5f * (float)(excl + pre[i]);
  // trapezoid node weights over z[0..S-2]: w_k = (dz_{k-1} [k >= 1] + dz_k [k <= S-3]) / 2, 0 at k = S - 1
  const float up = __shfl_up_sync(kFull, r.dzn[PER - 1], 1);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    r.dzp[i] = i > 0 ? r.dzn[i > 0 ? i - 1 : 0] : (lane > 0 ? up : 0.f);      // 0 at sample 0
    const bool s1 = lane == 31 && i == PER - 1, s2 = lane == 31 && i == PER - 2;   // samples S - 1, S - 2
    const float wk = s1 ? 0.f : 0.5f * (r.dzp[i] + (s2 ? 0.f : r.dzn[i]));
    r.wr[i] = in[i] ? wk * fmul(r.rho[i], r.rho[i]) : 0.f;                    // :263
  }
}

// v[0..7] summed over the warp: lane l returns the total of v[(l >> 2) & 7] (9 shuffles: the values are handed down a
// halving tree over lane bits 4, 3, 2, then two butterfly steps over bits 1, 0)
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
  float w[4], u[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
    w[i] = keep + __shfl_xor_sync(kFull, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? w[i] : w[i + 2], keep = b3 ? w[i + 2] : w[i];
    u[i] = keep + __shfl_xor_sync(kFull, send, 8);
  }
  const float send = b2 ? u[0] : u[1], keep = b2 ? u[1] : u[0];
  float t = keep + __shfl_xor_sync(kFull, send, 4);
  t += __shfl_xor_sync(kFull, t, 2);
  t += __shfl_xor_sync(kFull, t, 1);
  return t;
}

template <int PER>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_dt_fwd_blk_kernel(const float2 *__restrict__ inf, const float *__restrict__ z,
                                const float *__restrict__ wavelengths, int64_t N, int C,
                                const float *__restrict__ log_abs, const float *__restrict__ vol_c,
                                const float *__restrict__ table_x, const float *__restrict__ table_y, float F,
                                float *__restrict__ image, float *__restrict__ weights, float *__restrict__ regq) {
  constexpr int S = 32 * PER;
  __shared__ DtTables tabs;
  const DtTables *tab = &tabs;
  dt_load_tables(&tabs, table_x, table_y, log_abs);
  const bool uniform = dt_grid_is_uniform(tab);
  const uint32_t x_s = (uint32_t)__cvta_generic_to_shared(&tabs.x[0]);
  const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(&tabs.ys[0][0]);
  const uint32_t kappa_s = (uint32_t)__cvta_generic_to_shared(&tabs.kappa[0]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float vc = vol_c[0];
  const int64_t stride = (int64_t)gridDim.x * kRayWarps;
  for (int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp; ray < N; ray += stride) {
    float zz[PER];
    float2 v[PER];
    ld_blk<PER>(z + ray * S, lane, zz);
    ld_blk2<PER>(inf + ray * S, lane, v);
    const int k_lane = lane < C ? dt_channel_sel(__ldg(wavelengths + ray * C + lane)) : -1;   // lane c: table row of channel c
    const float nk_lane = k_lane >= 0 ? -lds_f32(kappa_s + 4 * k_lane) * 1.4426950408889634f : 0.f;  // exp(-kappa B/2) = 2^(nk B/2)
    float q[PER];
    double qsum = 0.0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { q[i] = fmaxf(v[i].x, 0.f); qsum += (double)q[i]; }
    st_blk<PER>(regq + ray * S, lane, q);                                                       // :271
    const float den = fadd((float)warp_sum(qsum), 1e-10f);
#pragma unroll
    for (int i = 0; i < PER; ++i) q[i] = fdiv(q[i], den);                                       // :268-269
    st_blk<PER>(weights + ray * S, lane, q);
    DtBlkRay<PER> r;
    dt_blk_setup<PER>(r, tab, x_s, ys_s, uniform, zz, v, lane);
    float part[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) part[c] = 0.f;
    // Channels in table order (wavelength c is row c of the response table wherever it is present - every shipped config
    // lists them that way, the STEREO mask only blanks entries): the row offset is a compile-time immediate of the load.
    const unsigned present = __ballot_sync(kFull, k_lane >= 0);
    const bool in_order = __all_sync(kFull, k_lane < 0 || k_lane == lane);
    if (in_order) {
      auto row = [&](auto row_tag, float nk) -> float {
        constexpr uint32_t OFF = (uint32_t)decltype(row_tag)::value * kYsRow;
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          const float2 ys = lds_f32x2_imm<OFF>(r.ysa[e]);
          const float R = fmaf(r.dxq[e], ys.y, ys.x);                           // :248 linear interpolation
          acc = fmaf(r.wr[e], ex2_ftz(nk * r.Bh[e]) * R, acc);                  // :264-265
        }
        return acc;
      };
      // an absent channel keeps response and absorption 0 (:243, :251): the pixel is 0 * vol_c * F
#define SNF_DT_ROW(c) if (present & (1u << c)) part[c] = row(std::integral_constant<int, c>{}, __shfl_sync(kFull, nk_lane, c))
      SNF_DT_ROW(0); SNF_DT_ROW(1); SNF_DT_ROW(2); SNF_DT_ROW(3); SNF_DT_ROW(4); SNF_DT_ROW(5); SNF_DT_ROW(6);
#undef SNF_DT_ROW
    } else {
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float acc = 0.f;
        if (c < C) {
          const int k = __shfl_sync(kFull, k_lane, c);
          const float nk = __shfl_sync(kFull, nk_lane, c);
          if (k >= 0) {
            const uint32_t koff = (uint32_t)k * kYsRow;
#pragma unroll
            for (int e = 0; e < PER; ++e) {
              const float2 ys = lds_f32x2(r.ysa[e] + koff);
              const float R = fmaf(r.dxq[e], ys.y, ys.x);
              acc = fmaf(r.wr[e], ex2_ftz(nk * r.Bh[e]) * R, acc);
            }
          }
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) part[cc] = cc == c ? acc : part[cc];     // static register indices
      }
    }
    const float J = warp_sum8(part, lane);
    const int ch = (lane >> 2) & 7;
    if ((lane & 3) == 0 && ch < C) image[ray * C + ch] = fmul(fmul(J, vc), F);
  }
}

template <int PER>
__global__ void __launch_bounds__(kRayWarps * 32)
    composite_dt_bwd_blk_kernel(const float2 *__restrict__ inf, const float *__restrict__ z,
                                const float *__restrict__ wavelengths, int64_t N, int C,
                                const float *__restrict__ log_abs, const float *__restrict__ vol_c,
                                const float *__restrict__ table_x, const float *__restrict__ table_y, float F,
                                const float *__restrict__ g_image, const float *__restrict__ g_regq,
                                float2 *__restrict__ g_inf, float *__restrict__ g_log_abs, float *__restrict__ g_vol_c) {
  constexpr int S = 32 * PER;
  __shared__ DtTables tabs;
  __shared__ float blk_acc[8];   // 7 kappa grads + vol_c grad of this CTA's rays
  const DtTables *tab = &tabs;
  if (threadIdx.x < 8) blk_acc[threadIdx.x] = 0.f;
  dt_load_tables(&tabs, table_x, table_y, log_abs);
  const bool uniform = dt_grid_is_uniform(tab);
  const uint32_t x_s = (uint32_t)__cvta_generic_to_shared(&tabs.x[0]);
  const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(&tabs.ys[0][0]);
  const uint32_t kappa_s = (uint32_t)__cvta_generic_to_shared(&tabs.kappa[0]);
  const uint32_t kon_s = (uint32_t)__cvta_generic_to_shared(&tabs.kappa_on[0]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float vc = vol_c[0];
  float gvc = 0.f;               // per-lane partial of dL/dvol_c, reduced once at the end
  const int64_t stride = (int64_t)gridDim.x * kRayWarps;
  for (int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp; ray < N; ray += stride) {
    float zz[PER], greg[PER];
    float2 v[PER];
    ld_blk<PER>(z + ray * S, lane, zz);
    ld_blk2<PER>(inf + ray * S, lane, v);
    if (g_regq != nullptr) {
      ld_blk<PER>(g_regq + ray * S, lane, greg);
    } else {
#pragma unroll
      for (int i = 0; i < PER; ++i) greg[i] = 0.f;
    }
    const int k_lane = lane < C ? dt_channel_sel(__ldg(wavelengths + ray * C + lane)) : -1;
    const float gi_lane = lane < C ? __ldg(g_image + ray * C + lane) : 0.f;
    const float kap_lane = k_lane >= 0 ? lds_f32(kappa_s + 4 * k_lane) : 0.f;
    DtBlkRay<PER> r;
    dt_blk_setup<PER>(r, tab, x_s, ys_s, uniform, zz, v, lane);
    // per sample, summed over the channels: s1 = sum_c Gc u_c (u = w tau), dth = dL/dtheta, dB = dL/dB
    float s1[PER], dth[PER], dB[PER], dkc[8];
#pragma unroll
    for (int e = 0; e < PER; ++e) s1[e] = dth[e] = dB[e] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) dkc[c] = 0.f;
    // one channel's contribution; ROW >= 0: table row known at compile time (its offset is an immediate of the load)
    auto channel = [&](auto row_tag, uint32_t koff, float kap, float gi) -> float {
      constexpr int ROW = decltype(row_tag)::value;
      const float nk = -kap * 1.4426950408889634f;
      const float Gc = gi * vc * F;            // dL/dJ with I = J * vol_c * F
      const float hk = -0.5f * kap * Gc;       // dL/dB_k = hk u_k  (A = kappa B / 2, dL/dA_k = -Gc u_k)
      float part = 0.f, dk = 0.f;
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        float2 ys;
        if constexpr (ROW >= 0) ys = lds_f32x2_imm<(uint32_t)(ROW >= 0 ? ROW : 0) * kYsRow>(r.ysa[e]);
        else ys = lds_f32x2(r.ysa[e] + koff);
        const float R = fmaf(r.dxq[e], ys.y, ys.x);
        const float eA = ex2_ftz(nk * r.Bh[e]);
        const float u = r.wr[e] * (eA * R);    // w_k tau_k (0 outside the table and beyond the last trapezoid node)
        part += u;
        dB[e] = fmaf(hk, u, dB[e]);
        dk = fmaf(r.Bh[e], u, dk);             // dL/dkappa = -Gc sum_k B_k / 2 u_k
        s1[e] = fmaf(Gc, u, s1[e]);            // dL/drho_k (emission part) = 2 / rho_k sum_c Gc u_c
        dth[e] = fmaf((Gc * eA) * r.wr[e], ys.y, dth[e]);   // dL/dtheta_k = sum_c Gc w_k rho_k^2 exp(-A) slope
      }
      gvc = fmaf(gi * F, part, gvc);
      return -Gc * dk;
    };
    // channels in table order wherever present (every shipped config; the STEREO mask only blanks entries): see the forward
    const unsigned present = __ballot_sync(kFull, k_lane >= 0);
    const bool in_order = __all_sync(kFull, k_lane < 0 || k_lane == lane);
    if (in_order) {
      // an absent channel's pixel is the constant 0: no gradient to anything but vol_c (0 * F)
#define SNF_DT_ROW(c) \
      if (present & (1u << c)) dkc[c] = channel(std::integral_constant<int, c>{}, 0u, __shfl_sync(kFull, kap_lane, c), __shfl_sync(kFull, gi_lane, c))
      SNF_DT_ROW(0); SNF_DT_ROW(1); SNF_DT_ROW(2); SNF_DT_ROW(3); SNF_DT_ROW(4); SNF_DT_ROW(5); SNF_DT_ROW(6);
#undef SNF_DT_ROW
    } else {
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float dk = 0.f;
        if (c < C) {
          const int k = __shfl_sync(kFull, k_lane, c);
          const float kap = __shfl_sync(kFull, kap_lane, c), gi = __shfl_sync(kFull, gi_lane, c);
          if (k >= 0) dk = channel(std::integral_constant<int, -1>{}, (uint32_t)k * kYsRow, kap, gi);
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) dkc[cc] = cc == c ? dk : dkc[cc];          // static register indices
      }
    }
    // kappa gradients of this ray: lane l gets channel (l >> 2) & 7 summed over the warp, then one shared atomic per channel
    {
      const float t = warp_sum8(dkc, lane);
      const int ch = (lane >> 2) & 7;
      const int k = __shfl_sync(kFull, k_lane, ch);
      if ((lane & 3) == 0 && k >= 0 && t != 0.f && lds_f32(kon_s + 4 * k) != 0.f) atomicAdd(&blk_acc[k], t);
    }
    // G_i = sum_{k>=i} dL/dB_k = dL/d term_i, term_i = dz_i (rho_i + rho_{i+1}): suffix sums inside the lane, one suffix scan
    // of the lanes' totals for all channels
    float G[PER];
    float run = 0.f;
#pragma unroll
    for (int i = PER - 1; i >= 0; --i) { run += dB[i]; G[i] = run; }
    float suf = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float n = __shfl_down_sync(kFull, suf, d);
      if (lane + d < 32) suf += n;
    }
    const float later = suf - run;                 // lanes above this one
#pragma unroll
    for (int i = 0; i < PER; ++i) G[i] += later;
    const float gup = __shfl_up_sync(kFull, G[PER - 1], 1);
    float2 o[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float Gp = i > 0 ? G[i > 0 ? i - 1 : 0] : (lane > 0 ? gup : 0.f);   // G of the previous sample (none at sample 0: dzp = 0)
      // d term_i / d rho_j: dz_j for i = j (dzn is 0 at S-1) and dz_{j-1} for i = j - 1
      const float dr = s1[i] * __fdividef(2.f, r.rho[i]) + r.dzn[i] * G[i] + r.dzp[i] * Gp;   // rho >= 1: the fast division is safe
      o[i].x = v[i].x > 0.f ? dr * r.rho[i] + greg[i] : 0.f;
      o[i].y = v[i].y > 0.f ? dth[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < PER / 2; ++i)
      __stcs(reinterpret_cast<float4 *>(g_inf + ray * S) + lane * (PER / 2) + i, make_float4(o[2 * i].x, o[2 * i].y, o[2 * i + 1].x, o[2 * i + 1].y));
  }
  gvc = warp_sum_f(gvc);
  if (lane == 0 && gvc != 0.f) atomicAdd(&blk_acc[7], gvc);
  __syncthreads();
  if (threadIdx.x < 7 && blk_acc[threadIdx.x] != 0.f) atomicAdd(&g_log_abs[threadIdx.x], blk_acc[threadIdx.x]);
  if (threadIdx.x == 7 && blk_acc[7] != 0.f) atomicAdd(g_vol_c, blk_acc[7]);
}

// ------------------------------------------------------------------------------------------ epilogue (a10)
__global__ void __launch_bounds__(kRayWarps * 32)
    render_epilogue_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                           const float *__restrict__ z, const float *__restrict__ weights,
                           const float *__restrict__ q, int64_t N, int S, float r0, int kind,
                           float *__restrict__ height_map, float *__restrict__ absorption_map,
                           float *__restrict__ reg, float gscale, float *__restrict__ g_q) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRayWarps + warp;
  if (ray >= N) return;
  const float o0 = rays_o[3 * ray], o1 = rays_o[3 * ray + 1], o2 = rays_o[3 * ray + 2];
  const float d0 = rays_d[3 * ray], d1 = rays_d[3 * ray + 1], d2 = rays_d[3 * ray + 2];
  double hsum = 0.0, asum = 0.0;
  for (int j = lane; j < S; j += 32) {
    const float zz = z[ray * S + j];
    const float p0 = fadd(o0, fmul(d0, zz)), p1 = fadd(o1, fmul(d1, zz)), p2 = fadd(o2, fmul(d2, zz));
    const float dist = __fsqrt_rn(sum3(fmul(p0, p0), fmul(p1, p1), fmul(p2, p2)));          // :101
    const float qq = q[ray * S + j];
    hsum += (double)fmul(weights[ray * S + j], dist);                                         // :102
    asum += (double)fsub(1.f, qq);                                                            // :99
    const float over = fmaxf(fsub(dist, r0), 0.f);
    float r, gq;
    if (kind == 0) { r = fmul(over, fsub(1.f, qq)); gq = -over * gscale; }                    // base_tracing.py:43-44
    else { r = fmul(over, fmaxf(qq, 0.f)); gq = qq > 0.f ? over * gscale : 0.f; }             // density_temperature.py:273-274
    reg[ray * S + j] = r;
    if (g_q != nullptr) g_q[ray * S + j] = gq;
  }
  hsum = warp_sum(hsum);
  asum = warp_sum(asum);
  if (lane == 0) { height_map[ray] = (float)hsum; absorption_map[ray] = (float)asum; }
}

// ------------------------------------------------------------------------------------------ SimpleStar (a7)
__global__ void __launch_bounds__(256) simple_star_kernel(const float4 *__restrict__ x, int64_t M, float rho_0,
                                                          float h0, float T0, float R_s, float t_ph,
                                                          float2 *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float4 p = x[i];
  const float r = __fsqrt_rn(sum3(fmul(p.x, p.x), fmul(p.y, p.y), fmul(p.z, p.z)));          // :73
  float rho, T;
  if (r <= 1.0f) {
    rho = rho_0;                                                                               // :83
    T = t_ph;                                                                                  // :90
  } else {
    rho = fmul(rho_0, expf(fmul(fdiv(1.f, h0), fsub(fdiv(1.f, r), 1.f))));                     // :85
    T = (r <= R_s) ? fadd(fmul(fsub(r, 1.f), fdiv(fsub(T0, t_ph), fsub(R_s, 1.f))), t_ph) : T0;   // :93, :96
  }
  out[i] = make_float2(logf(rho), log10f(T));                                                  // :86, :97
}

// ------------------------------------------------------------------------------------------ loss (a11)
__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e38f; }

__global__ void __launch_bounds__(256)
    train_loss_kernel(const float *__restrict__ coarse, const float *__restrict__ fine,
                      const float *__restrict__ target, const float *__restrict__ reg, int64_t n_img,
                      int64_t n_reg, int asinh_scaling, float a, float norm, float lambda_image,
                      float *__restrict__ acc /*[3]*/, float *__restrict__ g_coarse, float *__restrict__ g_fine,
                      int *__restrict__ finite_flag) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float sc = 0.f, sf = 0.f, sr = 0.f;
  bool bad = false;
  const float gcoef = 2.f * lambda_image / (float)n_img;   // d(lambda*mean((s-t)^2))/ds
  for (int64_t i = tid; i < n_img; i += stride) {
    const float c = coarse[i], f = fine[i], t = target[i];
    bad |= !finite_f(c) | !finite_f(f);
    float cs = c, fs = f, ts = t, dc = 1.f, df = 1.f;
    if (asinh_scaling) {   // ImageAsinhScaling: asinh(x/a)/asinh(1/a), vmax = 1
      cs = asinhf(c / a) / norm; fs = asinhf(f / a) / norm; ts = asinhf(t / a) / norm;
      dc = 1.f / (a * norm * sqrtf(1.f + (c / a) * (c / a)));
      df = 1.f / (a * norm * sqrtf(1.f + (f / a) * (f / a)));
    }
    const float ec = cs - ts, ef = fs - ts;
    sc += ec * ec; sf += ef * ef;
    g_coarse[i] = gcoef * ec * dc;
    g_fine[i] = gcoef * ef * df;
  }
  for (int64_t i = tid; i < n_reg; i += stride) {
    const float r = reg[i];
    bad |= !finite_f(r);
    sr += r;
  }
  // block-level reduction first: three atomics per block instead of three per warp on the same three addresses
  __shared__ float red[3][8];
  sc = warp_sum_f(sc); sf = warp_sum_f(sf); sr = warp_sum_f(sr);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = sc; red[1][w] = sf; red[2][w] = sr; }
  const bool any_bad = __syncthreads_or(bad);
  if (w == 0 && l < 3) {
    float t = 0.f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += red[l][i];
    atomicAdd(&acc[l], t);
  }
  if (any_bad && threadIdx.x == 0) atomicOr(finite_flag, 1);
}

// acc aliases losses+1 (no __restrict__): all three sums are read before anything is written
__global__ void train_loss_finalize_kernel(const float *acc, int64_t n_img, int64_t n_reg, float lambda_image,
                                           float lambda_reg, float *losses) {
  const float lc = acc[0] / (float)n_img, lf = acc[1] / (float)n_img, lr = acc[2] / (float)n_reg;
  losses[0] = lambda_image * (lc + lf) + lambda_reg * lr;
  losses[1] = lc; losses[2] = lf; losses[3] = lr;
}

}  // namespace snf

using namespace snf;

extern "C" int snf_composite_emission_fwd(const float *raw, const float *z, const float *rays_d, int64_t N, int S,
                                          float *image, float *weights, float *absorption, void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(raw); SNF_CHECK_PTR(z); SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(image); SNF_CHECK_PTR(weights);
  SNF_CHECK_PTR(absorption); SNF_CHECK_ALIGN(raw, 8);
  if (N < 0 || S < 2) return SNF_E_ARG;
  if (S > 256) return SNF_E_SHAPE;
  const unsigned grid = (unsigned)ceil_div64(N, kRayWarps);
  const float2 *raw2 = reinterpret_cast<const float2 *>(raw);
  // pair layout for S = 64, 128, 192, 256 (rows are then 256-byte multiples: vector accesses need the bases aligned)
  if (S % 64 == 0 && ((reinterpret_cast<uintptr_t>(raw) & 15) | (reinterpret_cast<uintptr_t>(z) & 7) |
                      (reinterpret_cast<uintptr_t>(weights) & 7) | (reinterpret_cast<uintptr_t>(absorption) & 7)) == 0) {
#define SNF_LAUNCH_P(NC) \
  composite_emission_fwd_pair_kernel<NC><<<grid, kRayWarps * 32, 0, (cudaStream_t)stream>>>(raw2, z, rays_d, N, image, weights, absorption)
    switch (S / 64) { case 1: SNF_LAUNCH_P(1); break; case 2: SNF_LAUNCH_P(2); break; case 3: SNF_LAUNCH_P(3); break; default: SNF_LAUNCH_P(4); break; }
#undef SNF_LAUNCH_P
    count_launch();
    return launch_status();
  }
#define SNF_LAUNCH(NCH) \
  composite_emission_fwd_kernel<NCH><<<grid, kRayWarps * 32, 0, (cudaStream_t)stream>>>(raw2, z, rays_d, N, S, image, weights, absorption)
  switch ((S + 31) / 32) {
    case 1: SNF_LAUNCH(1); break; case 2: SNF_LAUNCH(2); break; case 3: SNF_LAUNCH(3); break; case 4: SNF_LAUNCH(4); break;
    case 5: SNF_LAUNCH(5); break; case 6: SNF_LAUNCH(6); break; case 7: SNF_LAUNCH(7); break; default: SNF_LAUNCH(8); break;
  }
#undef SNF_LAUNCH
  count_launch();
  return launch_status();
}

extern "C" int snf_composite_emission_bwd(const float *raw, const float *z, const float *rays_d, int64_t N, int S,
                                          const float *g_image, const float *g_absorption, float *g_raw,
                                          void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(raw); SNF_CHECK_PTR(z); SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(g_image); SNF_CHECK_PTR(g_raw);
  SNF_CHECK_ALIGN(raw, 8); SNF_CHECK_ALIGN(g_raw, 8);
  if (N < 0 || S < 2) return SNF_E_ARG;
  if (S > 256) return SNF_E_SHAPE;
  const unsigned grid = (unsigned)ceil_div64(N, kRayWarps);
  const float2 *raw2 = reinterpret_cast<const float2 *>(raw);
  float2 *g2 = reinterpret_cast<float2 *>(g_raw);
  if (S % 64 == 0 && ((reinterpret_cast<uintptr_t>(raw) & 15) | (reinterpret_cast<uintptr_t>(z) & 7) |
                      (reinterpret_cast<uintptr_t>(g_raw) & 15) | (reinterpret_cast<uintptr_t>(g_absorption) & 7)) == 0) {
#define SNF_LAUNCH_P(NC) \
  composite_emission_bwd_pair_kernel<NC><<<grid, kRayWarps * 32, 0, (cudaStream_t)stream>>>(raw2, z, rays_d, N, g_image, g_absorption, g2)
    switch (S / 64) { case 1: SNF_LAUNCH_P(1); break; case 2: SNF_LAUNCH_P(2); break; case 3: SNF_LAUNCH_P(3); break; default: SNF_LAUNCH_P(4); break; }
#undef SNF_LAUNCH_P
    count_launch();
    return launch_status();
  }
#define SNF_LAUNCH(NCH) \
  composite_emission_bwd_kernel<NCH><<<grid, kRayWarps * 32, 0, (cudaStream_t)stream>>>(raw2, z, rays_d, N, S, g_image, g_absorption, g2)
  switch ((S + 31) / 32) {
    case 1: SNF_LAUNCH(1); break; case 2: SNF_LAUNCH(2); break; case 3: SNF_LAUNCH(3); break; case 4: SNF_LAUNCH(4); break;
    case 5: SNF_LAUNCH(5); break; case 6: SNF_LAUNCH(6); break; case 7: SNF_LAUNCH(7); break; default: SNF_LAUNCH(8); break;
  }
#undef SNF_LAUNCH
  count_launch();
  return launch_status();
}

extern "C" int snf_composite_dt_fwd(const float *inferences, const float *z, const float *wavelengths, int64_t N,
                                    int S, int C, const float *log_abs, const float *vol_c, const float *table_x,
                                    const float *table_y, float F, float *image, float *weights, float *regq,
                                    void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(inferences); SNF_CHECK_PTR(z); SNF_CHECK_PTR(wavelengths); SNF_CHECK_PTR(log_abs);
  SNF_CHECK_PTR(vol_c); SNF_CHECK_PTR(table_x); SNF_CHECK_PTR(table_y); SNF_CHECK_PTR(image);
  SNF_CHECK_PTR(weights); SNF_CHECK_PTR(regq); SNF_CHECK_ALIGN(inferences, 8);
  if (N < 0 || S < 3 || C <= 0) return SNF_E_ARG;
  if (S > 256 || C > 8) return SNF_E_SHAPE;
  int64_t nblk = ceil_div64(N, kRayWarps);
  if (nblk > 148 * 16) nblk = 148 * 16;             // persistent CTAs: the tables are built once per CTA
  const float2 *inf2 = reinterpret_cast<const float2 *>(inferences);
  if (S % 64 == 0 && ((reinterpret_cast<uintptr_t>(inferences) & 15) | (reinterpret_cast<uintptr_t>(z) & 7) |
                      (reinterpret_cast<uintptr_t>(weights) & 7) | (reinterpret_cast<uintptr_t>(regq) & 7)) == 0) {
#define SNF_LAUNCH_P(PER) \
  composite_dt_fwd_blk_kernel<PER><<<(unsigned)nblk, kRayWarps * 32, 0, (cudaStream_t)stream>>>( \
      inf2, z, wavelengths, N, C, log_abs, vol_c, table_x, table_y, F, image, weights, regq)
    switch (S / 64) { case 1: SNF_LAUNCH_P(2); break; case 2: SNF_LAUNCH_P(4); break; case 3: SNF_LAUNCH_P(6); break; default: SNF_LAUNCH_P(8); break; }
#undef SNF_LAUNCH_P
    count_launch();
    return launch_status();
  }
#define SNF_LAUNCH(NCH) \
  composite_dt_fwd_kernel<NCH><<<(unsigned)nblk, kRayWarps * 32, 0, (cudaStream_t)stream>>>( \
      inf2, z, wavelengths, N, S, C, log_abs, vol_c, table_x, table_y, F, image, weights, regq)
  switch ((S + 31) / 32) {
    case 1: SNF_LAUNCH(1); break; case 2: SNF_LAUNCH(2); break; case 3: SNF_LAUNCH(3); break; case 4: SNF_LAUNCH(4); break;
    case 5: SNF_LAUNCH(5); break; case 6: SNF_LAUNCH(6); break; case 7: SNF_LAUNCH(7); break; default: SNF_LAUNCH(8); break;
  }
#undef SNF_LAUNCH
  count_launch();
  return launch_status();
}

extern "C" int snf_composite_dt_bwd(const float *inferences, const float *z, const float *wavelengths, int64_t N,
                                    int S, int C, const float *log_abs, const float *vol_c, const float *table_x,
                                    const float *table_y, float F, const float *g_image, const float *g_regq,
                                    float *g_inferences, float *g_log_abs, float *g_vol_c, void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(inferences); SNF_CHECK_PTR(z); SNF_CHECK_PTR(wavelengths); SNF_CHECK_PTR(log_abs);
  SNF_CHECK_PTR(vol_c); SNF_CHECK_PTR(table_x); SNF_CHECK_PTR(table_y); SNF_CHECK_PTR(g_image);
  SNF_CHECK_PTR(g_inferences); SNF_CHECK_PTR(g_log_abs); SNF_CHECK_PTR(g_vol_c);
  SNF_CHECK_ALIGN(inferences, 8); SNF_CHECK_ALIGN(g_inferences, 8);
  if (N < 0 || S < 3 || C <= 0) return SNF_E_ARG;
  if (S > 256 || C > 8) return SNF_E_SHAPE;
  int64_t nblk = ceil_div64(N, kRayWarps);
  if (nblk > 148 * 16) nblk = 148 * 16;
  const float2 *inf2 = reinterpret_cast<const float2 *>(inferences);
  float2 *g2 = reinterpret_cast<float2 *>(g_inferences);
  if (S % 64 == 0 && ((reinterpret_cast<uintptr_t>(inferences) & 15) | (reinterpret_cast<uintptr_t>(z) & 7) |
                      (reinterpret_cast<uintptr_t>(g_inferences) & 15) | (reinterpret_cast<uintptr_t>(g_regq) & 7)) == 0) {
#define SNF_LAUNCH_P(PER) \
  composite_dt_bwd_blk_kernel<PER><<<(unsigned)nblk, kRayWarps * 32, 0, (cudaStream_t)stream>>>( \
      inf2, z, wavelengths, N, C, log_abs, vol_c, table_x, table_y, F, g_image, g_regq, g2, g_log_abs, g_vol_c)
    switch (S / 64) { case 1: SNF_LAUNCH_P(2); break; case 2: SNF_LAUNCH_P(4); break; case 3: SNF_LAUNCH_P(6); break; default: SNF_LAUNCH_P(8); break; }
#undef SNF_LAUNCH_P
    count_launch();
    return launch_status();
  }
#define SNF_LAUNCH(NCH) \
  composite_dt_bwd_kernel<NCH><<<(unsigned)nblk, kRayWarps * 32, 0, (cudaStream_t)stream>>>( \
      inf2, z, wavelengths, N, S, C, log_abs, vol_c, table_x, table_y, F, g_image, g_regq, g2, g_log_abs, g_vol_c)
  switch ((S + 31) / 32) {
    case 1: SNF_LAUNCH(1); break; case 2: SNF_LAUNCH(2); break; case 3: SNF_LAUNCH(3); break; case 4: SNF_LAUNCH(4); break;
    case 5: SNF_LAUNCH(5); break; case 6: SNF_LAUNCH(6); break; case 7: SNF_LAUNCH(7); break; default: SNF_LAUNCH(8); break;
  }
#undef SNF_LAUNCH
  count_launch();
  return launch_status();
}

extern "C" int snf_render_epilogue(const float *rays_o, const float *rays_d, const float *z_comb,
                                   const float *weights, const float *q, int64_t N, int S, float r0, int kind,
                                   float *height_map, float *absorption_map, float *reg, float reg_grad_scale,
                                   float *g_q, void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(rays_o); SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(z_comb); SNF_CHECK_PTR(weights); SNF_CHECK_PTR(q);
  SNF_CHECK_PTR(height_map); SNF_CHECK_PTR(absorption_map); SNF_CHECK_PTR(reg);
  if (N < 0 || S <= 0 || (kind != 0 && kind != 1)) return SNF_E_ARG;
  render_epilogue_kernel<<<(unsigned)ceil_div64(N, kRayWarps), kRayWarps * 32, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, z_comb, weights, q, N, S, r0, kind, height_map, absorption_map, reg, reg_grad_scale, g_q);
  count_launch();
  return launch_status();
}

extern "C" int snf_simple_star_fwd(const float *x, int64_t M, float rho_0, float h0, float T0, float R_s,
                                   float t_photosphere, float *out, void *stream) {
  if (M == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(x); SNF_CHECK_PTR(out); SNF_CHECK_ALIGN(x, 16); SNF_CHECK_ALIGN(out, 8);
  if (M < 0) return SNF_E_ARG;
  simple_star_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4 *>(x), M, rho_0, h0, T0, R_s, t_photosphere, reinterpret_cast<float2 *>(out));
  count_launch();
  return launch_status();
}

extern "C" int snf_train_loss(const float *coarse, const float *fine, const float *target, const float *reg,
                              int64_t N, int C, int64_t n_reg, int asinh_scaling, float asinh_a,
                              float lambda_image, float lambda_reg, float *losses, float *g_coarse, float *g_fine,
                              int *finite_flag, void *stream) {
  SNF_CHECK_PTR(coarse); SNF_CHECK_PTR(fine); SNF_CHECK_PTR(target); SNF_CHECK_PTR(reg); SNF_CHECK_PTR(losses);
  SNF_CHECK_PTR(g_coarse); SNF_CHECK_PTR(g_fine); SNF_CHECK_PTR(finite_flag);
  if (N <= 0 || C <= 0 || n_reg <= 0) return SNF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // reentrant without extra scratch: the raw sums accumulate in losses[1..3] and are finalised in place
  cudaError_t e = cudaMemsetAsync(losses, 0, 4 * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int64_t n_img = N * C;
  const float norm = (float)asinh(1.0 / (double)asinh_a);   // scaling.py:21
  const int64_t work = n_img > n_reg ? n_img : n_reg;
  const unsigned blocks = (unsigned)(ceil_div64(work, 256) < 592 ? ceil_div64(work, 256) : 592);
  train_loss_kernel<<<blocks, 256, 0, st>>>(coarse, fine, target, reg, n_img, n_reg, asinh_scaling, asinh_a, norm,
                                            lambda_image, losses + 1, g_coarse, g_fine, finite_flag);
  train_loss_finalize_kernel<<<1, 1, 0, st>>>(losses + 1, n_img, n_reg, lambda_image, lambda_reg, losses);
  count_launch(2);
  return launch_status();
}
