// How many thread-block clusters of size 2 / 4 / 8 with ~220 KB of shared memory per CTA are co-resident on this GPU?
// (decides whether weight multicast over 4-CTA clusters can use all 148 SMs)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out) { extern __shared__ char s[]; if (out) out[0] = s[0]; }
int main() {
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs, 1, 1);
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d SMs  (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
