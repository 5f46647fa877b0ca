// Probe (not part of the product): where does TMA put a box whose inner dimension (32 B) is SMALLER than the swizzle
// span (64 B)?  Global: 16 rows x 16 halfs, value = 16-byte chunk id (row*2 + half).  The box is the whole array,
// SWIZZLE_64B.  Prints, for every 16-byte chunk of shared memory, the chunk id found there.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_swizzle_probe tma_swizzle_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t s0 = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) ((uint16_t *)(smem + (s0 - (uint32_t)__cvta_generic_to_shared(smem))))[i] = 999;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(512u) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s0), "l"(&tm), "r"(b), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
  }
  __syncthreads();
  const uint16_t *sp = (const uint16_t *)(smem + (s0 - (uint32_t)__cvta_generic_to_shared(smem)));
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = sp[i];
}

int main() {
  uint16_t h[256];
  for (int r = 0; r < 16; r++)
    for (int c = 0; c < 16; c++) h[r * 16 + c] = (uint16_t)(r * 2 + c / 8);
  uint16_t *d, *o;
  cudaMalloc(&d, sizeof(h));
  cudaMalloc(&o, 2 * sizeof(h));
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[2] = {16, 16};
  cuuint64_t strides[1] = {32};
  cuuint32_t box[2] = {16, 16};
  cuuint32_t es[2] = {1, 1};
  for (int mode = 0; mode < 2; mode++) {
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        mode ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %s encode rc %d\n", mode ? "64B" : "32B", (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(o, 0xff, 2 * sizeof(h));
    probe<<<1, 64, 4096>>>(tm, o);
    cudaError_t e = cudaDeviceSynchronize();
    uint16_t g[512];
    cudaMemcpy(g, o, sizeof(g), cudaMemcpyDeviceToHost);
    printf("  %s\n  smem 16-byte chunk -> global chunk id (row*2 + half):\n", cudaGetErrorString(e));
    for (int c = 0; c < 64; c++) printf("%s%3d", (c % 8 == 0) ? "\n   " : " ", (int)g[c * 8]);
    printf("\n");
  }
  return 0;
}
