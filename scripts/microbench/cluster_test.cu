// Experiment: which cluster launch configurations does this B200 accept? (not part of the product)
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
__global__ void __launch_bounds__(512, 1) k(int *out) {
  extern __shared__ char smem[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) out[blockIdx.x] = (int)r;
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  int *d; cudaMalloc(&d, 1024 * sizeof(int));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  for (int smem : {0, 100 * 1024, 200 * 1024, 214 * 1024, 226 * 1024})
    for (int threads : {128, 512})
      for (int grid : {2, 148}) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = -1;
        cudaError_t eo = cudaOccupancyMaxActiveClusters(&nclusters, k, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("smem %6d threads %3d grid %3d: maxActiveClusters %d (%s) launch %s sync %s\n", smem, threads, grid, nclusters,
               cudaGetErrorString(eo), cudaGetErrorString(e), cudaGetErrorString(e2));
        cudaGetLastError();
      }
  return 0;
}
