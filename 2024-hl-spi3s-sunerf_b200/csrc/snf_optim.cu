// Flat-buffer optimiser step: global grad-norm -> clip(0.5) -> Adam, two launches, no host
// sync.  Replaces Lightning's gradient_clip_val=0.5 (sunerf/run_emission.py:72) + torch.optim.Adam(lr=1e-4)
// (sunerf/model/sunerf.py:30-35) over rendering.parameters(), applied to the single flat fp32 buffer that the
// NCCL all-reduce also uses.
#include "snf_common.cuh"

namespace snf {

constexpr int kNormBlocks = 592;   // 148 SMs x 4

__global__ void __launch_bounds__(256) sumsq_kernel(const float4 *__restrict__ g, int64_t n4, const float *__restrict__ tail,
                                                    int ntail, float scale, float *__restrict__ partial) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = g[i];
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) { const float v = tail[threadIdx.x] * scale; s += v * v; }
  __shared__ float red[8];
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
    for (int d = 4; d > 0; d >>= 1) s += __shfl_xor_sync(0xffu, s, d);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, float clip, float scale,
                                                   const float *__restrict__ partial, int nparts, float *__restrict__ norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x < 32) {   // every block re-reduces the (<=592) partial sums in the same fixed order
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) t += (double)partial[i];
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(t);
      float coef = clip > 0.f ? clip / (norm + 1e-6f) : 1.f;   // torch.nn.utils.clip_grad_norm_
      if (coef > 1.f) coef = 1.f;
      s_coef = coef * scale;
      if (blockIdx.x == 0 && norm_out != nullptr) norm_out[0] = norm;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;         // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// Device-resident schedule for CUDA-graph replay: sched = {lr, step, gamma, lr_floor} (doubles).  The Adam kernel reads
// lr and step from it, sched_advance_kernel applies the reference's per-batch ExponentialLR rule afterwards
// (sunerf/model/sunerf.py:36-40: step the scheduler while lr > 5e-5).
__global__ void __launch_bounds__(256) adam_sched_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                         float *__restrict__ v, int64_t n, const double *__restrict__ sched,
                                                         float b1, float b2, float eps, float clip, float scale,
                                                         const float *__restrict__ partial, int nparts,
                                                         float *__restrict__ norm_out) {
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) t += (double)partial[i];
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(t);
      float coef = clip > 0.f ? clip / (norm + 1e-6f) : 1.f;
      if (coef > 1.f) coef = 1.f;
      s_coef = coef * scale;
      if (blockIdx.x == 0 && norm_out != nullptr) norm_out[0] = norm;
      const double lr = sched[0], step = sched[1];
      const float bc1 = 1.f - (float)pow((double)b1, step);
      s_bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
      s_step_size = (float)lr / bc1;
    }
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

__global__ void sched_advance_kernel(double *sched) {
  sched[1] += 1.0;
  if (sched[0] > sched[3]) sched[0] *= sched[2];
}

}  // namespace snf

using namespace snf;

extern "C" int snf_adam_step_sched(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n,
                                   double *sched, float beta1, float beta2, float eps, float clip_norm, float grad_scale,
                                   float *scratch, float *norm_out, void *stream) {
  SNF_CHECK_PTR(params); SNF_CHECK_PTR(grads); SNF_CHECK_PTR(exp_avg); SNF_CHECK_PTR(exp_avg_sq); SNF_CHECK_PTR(scratch);
  SNF_CHECK_PTR(sched); SNF_CHECK_ALIGN(grads, 16); SNF_CHECK_ALIGN(sched, 8);
  if (n <= 0) return SNF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  const int ntail = (int)(n - n4 * 4);
  int blocks = (int)(ceil_div64(n4 > 0 ? n4 : 1, 256) < kNormBlocks ? ceil_div64(n4 > 0 ? n4 : 1, 256) : kNormBlocks);
  sumsq_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(grads), n4, grads + n4 * 4, ntail, grad_scale, scratch);
  const int ablocks = (int)(ceil_div64(n, 256) < kNormBlocks ? ceil_div64(n, 256) : kNormBlocks);
  adam_sched_kernel<<<ablocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, sched, beta1, beta2, eps, clip_norm, grad_scale,
                                             scratch, blocks, norm_out);
  sched_advance_kernel<<<1, 1, 0, st>>>(sched);
  count_launch(3);
  return launch_status();
}

extern "C" int snf_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int64_t step, float clip_norm, float grad_scale,
                             float *scratch, float *norm_out, void *stream) {
  SNF_CHECK_PTR(params); SNF_CHECK_PTR(grads); SNF_CHECK_PTR(exp_avg); SNF_CHECK_PTR(exp_avg_sq); SNF_CHECK_PTR(scratch);
  SNF_CHECK_ALIGN(grads, 16);
  if (n <= 0 || step < 1) return SNF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  const int ntail = (int)(n - n4 * 4);
  int blocks = (int)(ceil_div64(n4 > 0 ? n4 : 1, 256) < kNormBlocks ? ceil_div64(n4 > 0 ? n4 : 1, 256) : kNormBlocks);
  sumsq_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(grads), n4, grads + n4 * 4, ntail, grad_scale, scratch);
  const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const int ablocks = (int)(ceil_div64(n, 256) < kNormBlocks ? ceil_div64(n, 256) : kNormBlocks);
  adam_kernel<<<ablocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, clip_norm,
                                       grad_scale, scratch, blocks, norm_out);
  count_launch(2);
  return launch_status();
}
