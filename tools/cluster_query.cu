// How many clusters of 2 / 4 / 8 CTAs (1 CTA per SM: 200 KB dynamic shared memory) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) p[0] = 1; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
