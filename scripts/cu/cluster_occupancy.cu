// How many clusters of a 1024-thread / 135 KB CTA can be co-resident on this GPU, per cluster size?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) p[0] = 1; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s SMs=%d\n", prop.name, prop.multiProcessorCount);
  for (int smem : {135168, 100000, 48000}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int threads : {1024, 512}) {
      for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("smem=%6d threads=%4d cluster=%2d -> max active clusters %d (%s)\n", smem, threads, cs, n, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
