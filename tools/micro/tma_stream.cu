// Microbenchmark: per-SM L2 -> shared-memory streaming rate of cp.async.bulk through an mbarrier ring
// (all 148 SMs stream the same 3.7 MB buffer, as the MLP kernels do with the packed weights).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../2024-hl-spi3s-sunerf_b200/csrc/snf_tcgen05.cuh"
using namespace snf::tc;
__global__ void __launch_bounds__(64, 1) k(const uint8_t *src, int64_t src_bytes, int chunk, int nstage, int nchunks, long long *cyc, int lockstep) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bars = base + nstage * chunk;
  if (threadIdx.x == 0) { for (int s = 0; s < nstage; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (nstage + s), 1); } fence_barrier_init(); }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0; int64_t off = lockstep ? (int64_t)(blockIdx.x & 1) * chunk : (int64_t)blockIdx.x * chunk % src_bytes;
    for (int i = 0; i < nchunks; ++i) {
      mbar_wait(bars + 8 * (nstage + s), ph ^ 1);
      mbar_arrive_expect_tx(bars + 8 * s, chunk);
      bulk_g2s(base + s * chunk, src + off, chunk, bars + 8 * s);
      off += lockstep ? 2 * chunk : chunk; if (off + 2 * chunk > src_bytes) off = lockstep ? (int64_t)(blockIdx.x & 1) * chunk : 0;
      if (++s == nstage) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < nchunks; ++i) {
      mbar_wait(bars + 8 * s, ph);
      mbar_arrive(bars + 8 * (nstage + s));
      if (++s == nstage) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main() {
  const int64_t src_bytes = 3800 * 1024;
  uint8_t *src; long long *cyc; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int chunks[] = {16384, 32768};
  for (int lockstep = 0; lockstep < 2; ++lockstep)
  for (int c : chunks)
    for (int nstage = 3; nstage * c <= 192 * 1024 && nstage <= 8; ++nstage) {
      const int nchunks = (64 << 20) / c / 4;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 64, nstage * c + 256>>>(src, src_bytes, c, nstage, nchunks, cyc, lockstep);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("%s chunk %5d B x %d stages: %.1f B/clk/SM, %.0f clk/chunk  (%s)\n", lockstep ? "lockstep " : "staggered", c, nstage, (double)nchunks * c / avg, avg / nchunks, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
