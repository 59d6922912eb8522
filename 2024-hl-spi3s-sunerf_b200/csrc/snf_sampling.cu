// Ray-sample placement kernels (HBM-bound, fp32 with the reference's exact operation order).
//   K1 stratified_kernel : StratifiedSampler.forward      sunerf/train/sampling.py:68-102
//                          SphericalSampler.forward       sunerf/train/sampling.py:16-54 (same kernel, other bin ends)
//   K2 hier_kernel       : HierarchicalSampler.forward    sunerf/train/sampling.py:111-169
//   make_query_kernel    : o + d*z, cat time              sampling.py:100, base_tracing.py:64-65
#include "snf_common.cuh"

namespace snf {

unsigned long long g_launches = 0;
DeviceState g_devices[kMaxDevices];

// ------------------------------------------------------------------------------------------------
// K1: one warp per 32 rays.  Lane r computes the shell entry/exit of ray r once (two square roots and a
// division), then the warp walks over its rays: the scalars are broadcast by shuffle and lane <-> sample,
// so every access of the [N,S] streams is a full coalesced row and no per-ray work is repeated per sample.
// ------------------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(256) stratified_kernel(const float *__restrict__ rays_o,
                                                         const float *__restrict__ rays_d,
                                                         const float *__restrict__ t_vals,
                                                         const float *__restrict__ t_rand, int64_t N, int S,
                                                         float D, float solar_R, float *__restrict__ z_out,
                                                         float *__restrict__ pts_out, int rpw /* rays per warp, <= 32 */,
                                                         int spherical) {
  const int lane = threadIdx.x & 31;
  const int64_t ray0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw;
  if (ray0 >= N) return;
  const int64_t i = (lane < rpw && ray0 + lane < N) ? ray0 + lane : ray0;
  const float o0 = rays_o[3 * i], o1 = rays_o[3 * i + 1], o2 = rays_o[3 * i + 2];
  const float d0 = rays_d[3 * i], d1 = rays_d[3 * i + 1], d2 = rays_d[3 * i + 2];
  const float osq = sum3(fmul(o0, o0), fmul(o1, o1), fmul(o2, o2));
  const float r_obs = __fsqrt_rn(osq);                                            // :74
  const float qa = sum3(fmul(d0, d0), fmul(d1, d1), fmul(d2, d2));                // :77
  const float qb = sum3(fmul(fmul(2.f, o0), d0), fmul(fmul(2.f, o1), d1), fmul(fmul(2.f, o2), d2));  // :78
  const float qc = fsub(osq, fmul(solar_R, solar_R));                             // :80
  const float disc = fsub(fmul(qb, qb), fmul(fmul(4.f, qa), qc));
  const float hit = fdiv(fsub(-qb, __fsqrt_rn(disc)), fmul(2.f, qa));             // :81 (NaN on a miss)
  float my_near = fsub(r_obs, D);                                                 // :83
  float my_far = fadd(r_obs, D);                                                  // :84
  if (spherical) {   // SphericalSampler: entry / exit of the sphere of radius D (sampling.py:26-30; NaN when the ray misses it)
    const float qcd = fsub(osq, fmul(D, D));
    const float rt = __fsqrt_rn(fsub(fmul(qb, qb), fmul(fmul(4.f, qa), qcd)));
    my_near = fdiv(fsub(-qb, rt), fmul(2.f, qa));
    my_far = fdiv(fadd(-qb, rt), fmul(2.f, qa));
  }
  if (hit == hit) my_far = hit;                                                   // :87-88 / :37-38
  const int nr = (int)(N - ray0 < rpw ? N - ray0 : rpw);
  // S == 32 NJ: this lane's bin positions t[j-1], t[j], t[j+1] and their complements are the same for every ray of the
  // warp - loaded once, not once per ray (a fifth of the instructions of a ray)
  float tv[NJ > 0 ? NJ : 1][3], tc[NJ > 0 ? NJ : 1][3];
  if constexpr (NJ > 0) {
#pragma unroll
    for (int k = 0; k < NJ; ++k)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int j = min(max(k * 32 + lane + q - 1, 0), S - 1);
        tv[k][q] = __ldg(t_vals + j);
        tc[k][q] = fsub(1.f, tv[k][q]);
      }
  }
  auto do_ray = [&](int r, const float (&tr)[NJ > 0 ? NJ : 1]) {
    const float z_near = __shfl_sync(kFull, my_near, r), z_far = __shfl_sync(kFull, my_far, r);
    const float p0 = __shfl_sync(kFull, o0, r), p1 = __shfl_sync(kFull, o1, r), p2 = __shfl_sync(kFull, o2, r);
    const float e0 = __shfl_sync(kFull, d0, r), e1 = __shfl_sync(kFull, d1, r), e2 = __shfl_sync(kFull, d2, r);
    auto bin_edge = [&](int k) {                                                  // :90
      const float t = __ldg(t_vals + k);
      return fadd(fmul(z_near, fsub(1.f, t)), fmul(z_far, t));
    };
    const int64_t row = (ray0 + r) * S;
    auto finish = [&](int j, float z) {
      __stcs(z_out + row + j, z);
      if (pts_out != nullptr) {                                                   // :100
        pts_out[3 * (row + j)] = fadd(p0, fmul(e0, z));
        pts_out[3 * (row + j) + 1] = fadd(p1, fmul(e1, z));
        pts_out[3 * (row + j) + 2] = fadd(p2, fmul(e2, z));
      }
    };
    auto sample = [&](int j, float trj) {
      float z = bin_edge(j);
      if (t_rand != nullptr) {                                                    // :93-98
        const float hi = (j < S - 1) ? fmul(.5f, fadd(bin_edge(j + 1), z)) : z;
        const float lo = (j > 0) ? fmul(.5f, fadd(z, bin_edge(j - 1))) : z;
        z = fadd(lo, fmul(fsub(hi, lo), trj));
      }
      finish(j, z);
    };
    if constexpr (NJ > 0) {
#pragma unroll
      for (int k = 0; k < NJ; ++k) {
        const int j = k * 32 + lane;
        float z = fadd(fmul(z_near, tc[k][1]), fmul(z_far, tv[k][1]));            // :90
        if (t_rand != nullptr) {                                                  // :93-98
          const float zn = fadd(fmul(z_near, tc[k][2]), fmul(z_far, tv[k][2]));
          const float zp = fadd(fmul(z_near, tc[k][0]), fmul(z_far, tv[k][0]));
          const float hi = (j < S - 1) ? fmul(.5f, fadd(zn, z)) : z;
          const float lo = (j > 0) ? fmul(.5f, fadd(z, zp)) : z;
          z = fadd(lo, fmul(fsub(hi, lo), tr[k]));
        }
        finish(j, z);
      }
    } else {
      for (int j = lane; j < S; j += 32) sample(j, t_rand != nullptr ? __ldcs(t_rand + row + j) : 0.f);
    }
  };
  if constexpr (NJ > 0) {
    // S == 32 NJ: the jitter of RB rays is loaded before any of them is processed (RB x NJ x 128 B in flight per warp)
    constexpr int RB = 8;
    for (int r0 = 0; r0 < nr; r0 += RB) {
      float tr[RB][NJ > 0 ? NJ : 1];
#pragma unroll
      for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int k = 0; k < NJ; ++k)
          tr[b][k] = (t_rand != nullptr && r0 + b < nr) ? __ldcs(t_rand + (ray0 + r0 + b) * S + k * 32 + lane) : 0.f;
#pragma unroll
      for (int b = 0; b < RB; ++b)
        if (r0 + b < nr) do_ray(r0 + b, tr[b]);
    }
  } else {
    const float none[NJ > 0 ? NJ : 1] = {};
    for (int r = 0; r < nr; ++r) do_ray(r, none);
  }
}

// ------------------------------------------------------------------------------------------------
// K2: one warp per ray.  The ray's z, bin centres, CDF and the new samples live in the warp's slice of
// shared memory; the CDF is a warp scan in double (bit-equal to torch's CPU cumsum, which accumulates in
// double: every partial sum of these 62 addends is exactly representable, so association is irrelevant).
// The three searches of the reference (searchsorted of u in the CDF, and the two rank searches of the
// 64 + 128 merge that torch.sort performs) are O(1) per element instead of a binary search each:
//   first[j] = #{k : u_k < cdf_j}        arithmetic guess on the uniform u grid, corrected against the real u values
//   inds[k]  = #{j : cdf_j <= u_k} = #{j : first[j] <= k}                    -> histogram of first[] + prefix sum
//   a[j]     = #{k : new_z_k < z_j}      the samples of bin j are k in [first[j-1], first[j]): short search there,
//                                        corrected against the real new_z values
//   b[k]     = #{j : z_j <= new_z_k} = #{j : a[j] <= k}                      -> histogram of a[] + prefix sum
// Every count is verified against the actual values, so the result equals searchsorted / torch.sort exactly for any
// ascending u; a sortedness check falls back to an odd-even transposition sort (NaNs, rounding at a bin edge).
// ------------------------------------------------------------------------------------------------
// exclusive... inclusive prefix sum over cnt[0..n) -> out[k] = sum_{f<=k} cnt[f]; lane owns PER consecutive entries
template <int PER>
__device__ __forceinline__ void warp_prefix_counts(const int *cnt, int n, int lane, int (&out)[PER]) {
  int run = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane * PER + i;
    run += k < n ? cnt[k] : 0;
    out[i] = run;
  }
  int incl = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += v;
  }
  const int excl = incl - run;
#pragma unroll
  for (int i = 0; i < PER; ++i) out[i] += excl;
}

// PER = ceil(n_new / 32): new samples per lane (consecutive k).  SC, NC > 0: the sample counts as compile-time constants
// (the configs' 64 -> +128 on the shared u grid): every loop below has a known trip count and unrolls, the bounds tests
// and most of the index arithmetic fold away - the kernel is bound by instruction issue, a third of it was loop control.
#define SNF_UNROLL_IF_CT _Pragma("unroll (SC > 0 ? 8 : 1)")
template <int PER, int SC = 0, int NC = 0, int UPR = -1>
__global__ void __launch_bounds__(128) hier_kernel(const float *__restrict__ z_vals,
                                                   const float *__restrict__ weights,
                                                   const float *__restrict__ u, const float *__restrict__ cdf_in,
                                                   int64_t N, int S_rt, int n_new_rt, float *__restrict__ new_z,
                                                   float *__restrict__ z_comb, int64_t *__restrict__ inds,
                                                   float *__restrict__ cdf_out, int u_per_ray_rt) {
  const int S = SC > 0 ? SC : S_rt, n_new = NC > 0 ? NC : n_new_rt;
  const int u_per_ray = UPR >= 0 ? UPR : u_per_ray_rt;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int T = S + n_new;
  // CTA-wide: the u grid; per warp: zs[S] bins[S] cdf[S] nz[n_new] comb[T] | first[S] a[S] cnt[max(S, n_new) + 2]
  // (+ u_row[n_new] per warp when every ray has its own draws: HierarchicalSampler(perturb=True), sampling.py:144-146)
  float *us = sm;
  if (!u_per_ray)
    for (int k = threadIdx.x; k < n_new; k += blockDim.x) us[k] = u[k];
  __syncthreads();
  if (ray >= N) return;  // whole warp exits together; only __syncwarp below
  const int ncnt = (S > n_new ? S : n_new) + 2;
  const size_t wsz = (size_t)(3 * S + n_new + T + 2 * S + ncnt) + (u_per_ray ? n_new : 0);
  float *zs = sm + n_new + (size_t)warp * wsz;
  float *bins = zs + S, *cdf = bins + S, *nz = cdf + S, *comb = nz + n_new;
  int *first = reinterpret_cast<int *>(comb + T), *arank = first + S, *cnt = arank + S;
  const int nc = S - 1;  // CDF length == number of bin centres
  if (u_per_ray) {
    us = reinterpret_cast<float *>(cnt + ncnt);
    SNF_UNROLL_IF_CT
    for (int k = lane; k < n_new; k += 32) us[k] = __ldcs(u + ray * n_new + k);
  }

  SNF_UNROLL_IF_CT
  for (int j = lane; j < S; j += 32) zs[j] = __ldcs(z_vals + ray * S + j);
  SNF_UNROLL_IF_CT
  for (int i = lane; i < ncnt; i += 32) cnt[i] = 0;
  __syncwarp();
  SNF_UNROLL_IF_CT
  for (int j = lane; j < nc; j += 32) bins[j] = fmul(.5f, fadd(zs[j + 1], zs[j]));   // :118
  if (cdf_in != nullptr) {
    SNF_UNROLL_IF_CT
    for (int j = lane; j < nc; j += 32) cdf[j] = cdf_in[ray * nc + j];
  } else {
    const float *w = weights + ray * S + 1;  // weights[..., 1:-1]  :119
    const int nw = S - 2;
    double part = 0.0;
    SNF_UNROLL_IF_CT
    for (int j = lane; j < nw; j += 32) part += (double)fadd(__ldcs(w + j), 1e-5f);
    const float total = (float)warp_sum(part);                                       // :134
    double carry = 0.0;
    SNF_UNROLL_IF_CT
    for (int base = 0; base < nw; base += 32) {                                      // :137
      const int j = base + lane;
      const double p = (j < nw) ? (double)fdiv(fadd(w[j], 1e-5f), total) : 0.0;
      const double inc = warp_incl_sum(p, lane) + carry;
      if (j < nw) cdf[j + 1] = (float)inc;
      carry = __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) cdf[0] = 0.f;                                                     // :138
  }
  __syncwarp();
  if (cdf_out != nullptr)
    SNF_UNROLL_IF_CT
    for (int j = lane; j < nc; j += 32) cdf_out[ray * nc + j] = cdf[j];

  // ---- first[j] = #{k : u_k < cdf_j} and its histogram (ascending shared u only)
  if (!u_per_ray)
  SNF_UNROLL_IF_CT
  for (int j = lane; j < nc; j += 32) {
    const float c = cdf[j];
    int k0 = __float2int_ru(c * (float)(n_new - 1));        // exact on the ideal grid k / (n - 1); NaN -> 0
    k0 = min(max(k0, 0), n_new);
    while (k0 > 0 && us[k0 - 1] >= c) --k0;
    while (k0 < n_new && us[k0] < c) ++k0;
    first[j] = k0;
    atomicAdd(&cnt[k0], 1);
  }
  __syncwarp();
  // ---- inds[k] = #{j : cdf_j <= u_k} (searchsorted right=True, :149) and the inverse-CDF samples
  int ind[PER];
  if (!u_per_ray) {
    warp_prefix_counts<PER>(cnt, n_new, lane, ind);
  } else {   // unordered per-ray draws: one upper-bound binary search per draw, as torch.searchsorted(right=True) does
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = lane * PER + i;
      int lo = 0, hi = nc;
      if (k < n_new) {
        const float uu = us[k];
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
        }
      }
      ind[i] = lo;
    }
  }
  __syncwarp();
  SNF_UNROLL_IF_CT
  for (int i = lane; i < ncnt; i += 32) cnt[i] = 0;         // reused for the merge below
  float nzr[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane * PER + i;
    nzr[i] = 0.f;
    if (k < n_new) {
      const float uu = us[k];
      const int lo = ind[i];
      const int below = max(lo - 1, 0), above = min(lo, nc - 1);                     // :152-153
      const float cb = cdf[below], ca = cdf[above], bb = bins[below], ba = bins[above];
      float denom = fsub(ca, cb);                                                    // :164
      if (denom < 1e-5f) denom = 1.f;                                                // :165
      const float t = fdiv(fsub(uu, cb), denom);                                     // :166
      const float s = fadd(bb, fmul(t, fsub(ba, bb)));                               // :167
      nz[k] = s;
      nzr[i] = s;
      if (inds != nullptr) inds[ray * n_new + k] = lo;
    }
  }
  __syncwarp();
  bool sorted = true;
  if constexpr (NC > 0 && PER == 4) {
    // the lane's four consecutive samples leave as one 16-byte store and are checked for order in registers
    __stcs(reinterpret_cast<float4 *>(new_z + ray * n_new) + lane, make_float4(nzr[0], nzr[1], nzr[2], nzr[3]));
    const float nxt = __shfl_down_sync(kFull, nzr[0], 1);
    sorted = (nzr[0] <= nzr[1]) & (nzr[1] <= nzr[2]) & (nzr[2] <= nzr[3]) & (lane == 31 || nzr[3] <= nxt);
  } else {
    SNF_UNROLL_IF_CT
    for (int k = lane; k < n_new; k += 32) __stcs(new_z + ray * n_new + k, nz[k]);   // coalesced copy of the row
    SNF_UNROLL_IF_CT
    for (int k = lane; k < n_new - 1; k += 32) sorted &= (nz[k] <= nz[k + 1]);
  }

  // ---- sort(cat(z, new_z))  :123
  SNF_UNROLL_IF_CT
  for (int j = lane; j < S - 1; j += 32) sorted &= (zs[j] <= zs[j + 1]);
  sorted = __all_sync(kFull, sorted) && !u_per_ray;        // first[] exists for the shared ascending grid only
  if (sorted) {
    // a[j] = #{k : new_z_k < z_j}: the samples drawn from bin j (inds == j) are k in [first[j-1], first[j])
    SNF_UNROLL_IF_CT
    for (int j = lane; j < S; j += 32) {
      const float v = zs[j];
      int lo = j >= 1 ? first[min(j - 1, nc - 1)] : 0, hi = j < nc ? first[j] : n_new;
      if (hi < lo) hi = lo;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (nz[mid] < v) lo = mid + 1; else hi = mid;
      }
      while (lo > 0 && nz[lo - 1] >= v) --lo;               // exactness does not rest on the bracket
      while (lo < n_new && nz[lo] < v) ++lo;
      arank[j] = lo;
      comb[j + lo] = v;                                     // rank of z_j: j + #{new_z < z_j}
      atomicAdd(&cnt[lo], 1);
    }
    __syncwarp();
    int bk[PER];                                            // b[k] = #{j : z_j <= new_z_k} = #{j : a[j] <= k}
    warp_prefix_counts<PER>(cnt, n_new, lane, bk);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = lane * PER + i;
      if (k < n_new) comb[k + bk[i]] = nz[k];               // rank of new_z_k: k + #{z <= new_z_k}
    }
  } else {  // rare: an input is not monotone (rounding at a bin edge, NaN) -> odd-even transposition sort
    SNF_UNROLL_IF_CT
    for (int j = lane; j < S; j += 32) comb[j] = zs[j];
    SNF_UNROLL_IF_CT
    for (int k = lane; k < n_new; k += 32) comb[S + k] = nz[k];
    __syncwarp();
    for (int phase = 0; phase < T; ++phase) {
      for (int p = 2 * lane + (phase & 1); p + 1 < T; p += 64) {
        const float a = comb[p], b = comb[p + 1];
        if (!(a <= b) && (b == b)) { comb[p] = b; comb[p + 1] = a; }   // NaNs sink to the end like torch.sort
      }
      __syncwarp();
    }
  }
  __syncwarp();
  SNF_UNROLL_IF_CT
  for (int j = lane; j < T; j += 32) __stcs(z_comb + ray * T + j, comb[j]);
}

// query[n,s,:] = (o + d*z, t)
__global__ void __launch_bounds__(256) make_query_kernel(const float *__restrict__ rays_o,
                                                         const float *__restrict__ rays_d,
                                                         const float *__restrict__ z,
                                                         const float *__restrict__ times, int64_t N, int S,
                                                         float4 *__restrict__ query) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * S) return;
  const int64_t i = idx / S;
  const float zz = z[idx];
  float4 q;
  q.x = fadd(rays_o[3 * i], fmul(rays_d[3 * i], zz));
  q.y = fadd(rays_o[3 * i + 1], fmul(rays_d[3 * i + 1], zz));
  q.z = fadd(rays_o[3 * i + 2], fmul(rays_d[3 * i + 2], zz));
  q.w = times[i];
  query[idx] = q;
}

}  // namespace snf

using namespace snf;

extern "C" int snf_version(void) { return SNF_VERSION; }
extern "C" int64_t snf_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" void snf_count_launches(int64_t n) { count_launch((int)n); }   // kernels launched by a CUDA-graph replay
extern "C" const char *snf_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SNF_E_ARG: return "bad argument (null pointer or non-positive size)";
    case SNF_E_SHAPE: return "shape not supported by the compiled kernels";
    case SNF_E_ALIGN: return "pointer alignment";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown";
  }
}

// hier_kernel needs up to 51 KB of dynamic shared memory at its accepted limits (S = 256, n_new = 512): opt in per device
constexpr int kHierMaxSmem = 72 * 1024;
int snf_sampling_set_attributes() {
  cudaError_t e = cudaSuccess;
#define SNF_ATTR(PER) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hier_kernel<PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHierMaxSmem)
  SNF_ATTR(1); SNF_ATTR(2); SNF_ATTR(4); SNF_ATTR(8); SNF_ATTR(16);
#undef SNF_ATTR
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(hier_kernel<4, 64, 128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHierMaxSmem);
  return (int)e;
}

static int sample_bins(const float *rays_o, const float *rays_d, const float *t_vals, const float *t_rand, int64_t N, int S,
                       float distance, float solar_R, float *z_vals, float *points, int spherical, void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(rays_o); SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(t_vals); SNF_CHECK_PTR(z_vals);
  if (N < 0 || S <= 0) return SNF_E_ARG;
  int rpw = 32;                                   // small batches: fewer rays per warp so the grid still fills the GPU
  while (rpw > 1 && N / rpw < 148 * 16) rpw >>= 1;
  const unsigned grid = (unsigned)ceil_div64(N, 8 * rpw);
  if (S == 64)
    stratified_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_vals, t_rand, N, S, distance, solar_R, z_vals, points, rpw, spherical);
  else
    stratified_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_vals, t_rand, N, S, distance, solar_R, z_vals, points, rpw, spherical);
  count_launch();
  return launch_status();
}

extern "C" int snf_stratified_sample(const float *rays_o, const float *rays_d, const float *t_vals,
                                     const float *t_rand, int64_t N, int S, float distance, float solar_R,
                                     float *z_vals, float *points, void *stream) {
  return sample_bins(rays_o, rays_d, t_vals, t_rand, N, S, distance, solar_R, z_vals, points, 0, stream);
}

extern "C" int snf_spherical_sample(const float *rays_o, const float *rays_d, const float *t_vals,
                                    const float *t_rand, int64_t N, int S, float distance, float solar_R,
                                    float *z_vals, float *points, void *stream) {
  return sample_bins(rays_o, rays_d, t_vals, t_rand, N, S, distance, solar_R, z_vals, points, 1, stream);
}

static int hier_launch(const float *z_vals, const float *weights, const float *u, const float *cdf_in, int64_t N, int S,
                       int n_new, float *new_z, float *z_comb, int64_t *inds, float *cdf_out, int u_per_ray, void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(z_vals); SNF_CHECK_PTR(u); SNF_CHECK_PTR(new_z); SNF_CHECK_PTR(z_comb);
  if (weights == nullptr && cdf_in == nullptr) return SNF_E_ARG;
  if (N < 0 || S < 3 || n_new <= 0) return SNF_E_ARG;
  if (S > 256 || n_new > 512) return SNF_E_SHAPE;
  const int warps = 4;
  const int ncnt = (S > n_new ? S : n_new) + 2;
  const size_t smem = ((size_t)n_new + (size_t)warps * (3 * S + n_new + (S + n_new) + 2 * S + ncnt + (u_per_ray ? n_new : 0))) * sizeof(float);
  if (smem > (size_t)kHierMaxSmem) return SNF_E_SHAPE;
  if (smem > 48 * 1024)
    if (int e = snf_device_setup(nullptr)) return e;
  const unsigned grid = (unsigned)ceil_div64(N, warps);
#define SNF_LAUNCH(PER) \
  hier_kernel<PER><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(z_vals, weights, u, cdf_in, N, S, n_new, new_z, z_comb, inds, cdf_out, u_per_ray)
  const int per = (n_new + 31) / 32;
  if (S == 64 && n_new == 128 && !u_per_ray && (reinterpret_cast<uintptr_t>(new_z) & 15) == 0) {   // the configs' shape, compile-time sized
    hier_kernel<4, 64, 128, 0><<<grid, warps * 32, smem, (cudaStream_t)stream>>>(z_vals, weights, u, cdf_in, N, S, n_new, new_z, z_comb,
                                                                              inds, cdf_out, 0);
    count_launch();
    return launch_status();
  }
  if (per <= 1) SNF_LAUNCH(1); else if (per <= 2) SNF_LAUNCH(2); else if (per <= 4) SNF_LAUNCH(4);
  else if (per <= 8) SNF_LAUNCH(8); else SNF_LAUNCH(16);
#undef SNF_LAUNCH
  count_launch();
  return launch_status();
}

extern "C" int snf_hier_resample(const float *z_vals, const float *weights, const float *u, const float *cdf_in,
                                 int64_t N, int S, int n_new, float *new_z, float *z_comb, int64_t *inds,
                                 float *cdf_out, void *stream) {
  return hier_launch(z_vals, weights, u, cdf_in, N, S, n_new, new_z, z_comb, inds, cdf_out, 0, stream);
}

extern "C" int snf_hier_resample_perturb(const float *z_vals, const float *weights, const float *u_rand, const float *cdf_in,
                                         int64_t N, int S, int n_new, float *new_z, float *z_comb, int64_t *inds,
                                         float *cdf_out, void *stream) {
  return hier_launch(z_vals, weights, u_rand, cdf_in, N, S, n_new, new_z, z_comb, inds, cdf_out, 1, stream);
}

extern "C" int snf_make_query(const float *rays_o, const float *rays_d, const float *z, const float *times,
                              int64_t N, int S, float *query, void *stream) {
  if (N == 0) return 0;   // an empty batch is valid (and its tensors have null data pointers)
  SNF_CHECK_PTR(rays_o); SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(z); SNF_CHECK_PTR(times); SNF_CHECK_PTR(query);
  SNF_CHECK_ALIGN(query, 16);
  if (N < 0 || S <= 0) return SNF_E_ARG;
  make_query_kernel<<<(unsigned)ceil_div64(N * S, 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, z, times, N, S, reinterpret_cast<float4 *>(query));
  count_launch();
  return launch_status();
}
