// Microbenchmark: issue cost (cycles, single thread, back to back) of the synchronisation primitives the MLP kernels use.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../2024-hl-spi3s-sunerf_b200/csrc/snf_tcgen05.cuh"
using namespace snf::tc;
__device__ __forceinline__ void arrive_remote_relaxed(uint32_t a) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(long long *out, int which) {
  __shared__ __align__(16) uint64_t bars[8];
  __shared__ uint4 buf[128];
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(b0, (1 << 20) - 1); mbar_init(b1, 1); fence_barrier_init(); }
  cluster_sync_all();
  const int N = 256;
  if (rank == 1 && threadIdx.x == 0) {
    const uint32_t remote = mapa_shared(b0, 0);
    long long t;
    if (which == 0) { t = clock64(); for (int i = 0; i < N; ++i) mbar_arrive_remote(remote); out[0] = (clock64() - t) / N; }
    if (which == 1) { t = clock64(); for (int i = 0; i < N; ++i) arrive_remote_relaxed(remote); out[1] = (clock64() - t) / N; }
    if (which == 2) { t = clock64(); for (int i = 0; i < N; ++i) mbar_arrive(b0); out[2] = (clock64() - t) / N; }
    if (which == 3) { t = clock64(); for (int i = 0; i < N; ++i) fence_proxy_async_smem(); out[3] = (clock64() - t) / N; }
    if (which == 4) { t = clock64(); for (int i = 0; i < N; ++i) { buf[i & 127] = make_uint4(i, i, i, i); fence_proxy_async_smem(); } out[4] = (clock64() - t) / N; }
    if (which == 5) { t = clock64(); for (int i = 0; i < N; ++i) { buf[i & 127] = make_uint4(i, i, i, i); fence_proxy_async_smem(); mbar_arrive_remote(remote); } out[5] = (clock64() - t) / N; }
    if (which == 6) { t = clock64(); for (int i = 0; i < N; ++i) { buf[i & 127] = make_uint4(i, i, i, i); fence_proxy_async_smem(); arrive_remote_relaxed(remote); } out[6] = (clock64() - t) / N; }
    if (which == 7) { t = clock64(); for (int i = 0; i < N; ++i) tcgen05_fence_before(); out[7] = (clock64() - t) / N; }
    mbar_arrive(b1);
    if (which == 8) { t = clock64(); for (int i = 0; i < N; ++i) (void)mbar_try_wait(b1, 0); out[8] = (clock64() - t) / N; }
    if (which == 9) { t = clock64(); for (int i = 0; i < N; ++i) (void)mbar_try_wait_cluster(b1, 0); out[9] = (clock64() - t) / N; }
    if (which == 10) { t = clock64(); for (int i = 0; i < N; ++i) { buf[i & 127] = make_uint4(i, i, i, i); __syncwarp(); } out[10] = (clock64() - t) / N; }
    if (which == 11) { t = clock64(); for (int i = 0; i < N; ++i) { asm volatile("fence.acq_rel.cluster;" ::: "memory"); } out[11] = (clock64() - t) / N; }
  }
  cluster_sync_all();
}
#include <cstdlib>
int main(int argc, char **argv) {
  long long *d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
  int which = argc > 1 ? atoi(argv[1]) : 0;
  k<<<2, 128>>>(d, which); cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char *names[] = {"remote arrive release.cluster", "remote arrive relaxed.cluster", "local arrive", "fence.proxy.async", "st.shared + fence.proxy.async",
                         "st + fence.proxy + remote release", "st + fence.proxy + remote relaxed", "tcgen05.fence::before", "try_wait (complete, cta)", "try_wait (complete, acquire.cluster)",
                         "st.shared + syncwarp", "fence.acq_rel.cluster"};
  printf("%-40s %lld cycles\n", names[which], h[which]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
