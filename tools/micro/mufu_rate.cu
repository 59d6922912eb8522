// Microbenchmark: per-SM throughput of MUFU.SIN / MUFU.EX2 / polynomial sine (FMA pipe) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ float op(float x) {
  if (MODE == 0) return __sinf(x);
  if (MODE == 1) return exp2f(x) ;
  if (MODE == 2) return __cosf(x);
  if (MODE == 3) {   // polynomial sine, degree 7, range reduction by pi
    const float t = x * 0.318309886f;
    const float kf = t + 12582912.f;
    const float k = kf - 12582912.f;
    const float r = fmaf(k, -3.14159265f, x);
    const float r2 = r * r;
    float p = fmaf(-1.9515296e-4f, r2, 8.3321608e-3f);
    p = fmaf(p, r2, -1.6666654e-1f);
    const float rr = r * r2;
    const float s = fmaf(p, rr, r);
    return __uint_as_float(__float_as_uint(s) ^ (__float_as_uint(kf) << 31));
  }
  return x;
}
template <int MODE> __global__ void k(float *out, long long *cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<MODE>(a[i]) + 1.0f;
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name) {
  float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE><<<148, 1024>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-10s %.2f ops/clk/SM  (%.0f cycles)\n", name, 1024.0 * 8 * iters / avg, avg);
  cudaFree(out); cudaFree(cyc);
}
int main() { run<0>("sin.approx"); run<1>("ex2"); run<2>("cos.approx"); run<3>("poly sin"); run<4>("fadd only"); return 0; }
