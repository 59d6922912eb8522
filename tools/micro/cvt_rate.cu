// Microbenchmark: per-SM throughput of the fp32 -> bf16x2 pack (cvt.rn.bf16x2.f32 = F2FP) against a manual integer
// round-to-nearest-even pack, alone and mixed with MUFU.SIN (do they share the XU pipe?).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t pack_cvt(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t *>(&v); }
__device__ __forceinline__ uint32_t pack_int(float a, float b) {
  uint32_t x = __float_as_uint(a), y = __float_as_uint(b);
  x += 0x7FFFu + ((x >> 16) & 1u);
  y += 0x7FFFu + ((y >> 16) & 1u);
  return __byte_perm(x, y, 0x7632);
}
template <int MODE> __global__ void k(uint32_t *out, long long *cyc, int iters) {
  float a[8]; uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float x = a[i], y = a[i + 1];
      if (MODE == 2 || MODE == 3) { x = __sinf(x); y = __sinf(y); }
      uint32_t p = (MODE == 0 || MODE == 2) ? pack_cvt(x, y) : pack_int(x, y);
      acc ^= p;
      a[i] = x + 1.0f + __uint_as_float(p & 0x3f800000u) * 1e-30f; a[i + 1] = y + 2.0f;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(a[0]);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name) {
  uint32_t *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE><<<148, 1024>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-28s %.2f elements/clk/SM  (%.0f cycles)\n", name, 1024.0 * 8 * iters / avg, avg);
  cudaFree(out); cudaFree(cyc);
}
int main() { run<0>("cvt.rn.bf16x2 pack"); run<1>("integer RNE pack"); run<2>("sin.approx + cvt pack"); run<3>("sin.approx + integer pack"); return 0; }
