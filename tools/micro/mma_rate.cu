// Sustained tcgen05.mma rate of a CTA pair (cta_group::2, M = 256, bf16 -> fp32), all SMs busy:
//   N = 256 or 128 per instruction, A operand from shared memory (SS) or from TMEM (TS).
// Decides whether a ping-pong design with N = 128 instructions and a TMEM-resident activation operand keeps the tensor
// pipe at its N = 256 / shared-memory rate.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../2024-hl-spi3s-sunerf_b200/csrc/snf_tcgen05.cuh"
using namespace snf::tc;

constexpr int A_IMG = 2 * 128 * 128, B_IMG = 2 * 128 * 128;   // K = 128: 2 k-slabs of 128 rows x 128 B

struct Args { long long *cycles; int n, use_ts, reps, stage_sync; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_kernel(Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t rank = cluster_ctarank();
  const uint32_t sA = base, sB = base + A_IMG, sBar = sB + B_IMG, sSlot = sBar + 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (A_IMG + B_IMG) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(sBar, 1); mbar_init(sBar + 8, 1); mbar_init(sBar + 32, 1); mbar_init(sBar + 40, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc_2cta(sSlot, 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + A_IMG + B_IMG + 16);
  const uint32_t tm_row = tmem + ((uint32_t)(warp * 32) << 16);
  if (p.use_ts) {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = 0x3c003c00u;
    for (int kc = 0; kc < 4; ++kc) tmem_st16(tm_row + 256 + kc * 16, v);
    tmem_st_wait();
    tcgen05_fence_before();
  }
  __syncthreads();
  if (rank == 1 && tid == 0) mbar_arrive_remote(mapa_shared(sBar + 8, 0));
  if (rank == 0 && tid == 0) {
    mbar_wait_cluster(sBar + 8, 0);
    tcgen05_fence_after();
    const uint32_t idesc = idesc_bf16(256, p.n);
    // descriptors precomputed, eight instructions per trip fully unrolled, the whole warp-uniform path in registers:
    // the issue loop itself must not be what is measured
    uint64_t bd[8], ad[8];
    uint32_t at[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bd[j] = smem_desc(sB + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024);
      ad[j] = smem_desc(sA + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024);
      at[j] = tmem + 256 + j * 8;
    }
    const uint32_t d = tmem;   // one accumulator, as in a real K loop
    const long long t0 = clock64();
    if (p.stage_sync) {
      // what a real issuer does around every 8 instructions: wait for the weight stage (here: a barrier that is already
      // complete), fence, issue, commit the stage's "empty" barrier
      mbar_arrive(sBar + 32);                    // phase 0 of the dummy "full" barrier completes once and stays complete
#pragma unroll 1
      for (int r = 0; r < p.reps; ++r) {
        mbar_wait(sBar + 32, 0);
        tcgen05_fence_after();
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_ts_2cta(d, at[j], bd[j], idesc, 1);
        mma_commit_2cta(sBar + 40, 1);
      }
    } else if (p.use_ts) {
#pragma unroll 1
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_ts_2cta(d, at[j], bd[j], idesc, 1);
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_ss_2cta(d, ad[j], bd[j], idesc, 1);
      }
    }
    mma_commit_2cta(sBar, 3);
    mbar_wait(sBar, 0);
    p.cycles[blockIdx.x >> 1] = clock64() - t0;
  } else {
    mbar_wait(sBar, 0);
  }
  tcgen05_fence_after();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

int main() {
  const int smem = A_IMG + B_IMG + 64, reps = 4000, grid = 148;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *dc;
  cudaMalloc(&dc, sizeof(long long) * grid);
  for (int use_ts = 0; use_ts < 3; ++use_ts)
    for (int n : {256, 128, 64}) {
      Args p{dc, n, use_ts > 0, reps, use_ts == 2};
      for (int it = 0; it < 2; ++it) {
        rate_kernel<<<grid, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> c(grid / 2);
      cudaMemcpy(c.data(), dc, sizeof(long long) * grid / 2, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (auto v : c) mean += (double)v;
      mean /= c.size();
      const double per = mean / (reps * 8.0);
      const double flop_clk_sm = 2.0 * 256 * n * 16 / per / 2;
      printf("%s N=%3d: %.1f cycles per MMA, %.0f flop/clk/SM (%.0f %% of 8192)\n", use_ts == 2 ? "TMEM + stage sync" : use_ts ? "A from TMEM  " : "A from shared", n, per,
             flop_clk_sm, 100.0 * flop_clk_sm / 8192);
    }
  return 0;
}
