// Probe (not part of the product): cuTensorMapEncodeTiled elementStrides.  Global: 32 rows x 16 halfs (32 B rows),
// value = row.  Box {16, BOX1} with elementStrides {1, 2}, start row START: which rows land in shared memory, how
// many bytes does the transaction count, what happens at a negative start?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride_probe tma_stride_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t *out, int start, uint32_t tx) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t off = ((1024u - ((uint32_t)__cvta_generic_to_shared(smem) & 1023u)) & 1023u);
  uint16_t *sp = (uint16_t *)(smem + off);
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sp);
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sp[i] = 999;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(tx) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s0), "l"(&tm), "r"(b), "r"(0), "r"(start) : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 20000000LL)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    out[600] = (uint16_t)ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = sp[i];
}

int main() {
  uint16_t h[32 * 16];
  for (int r = 0; r < 32; r++)
    for (int c = 0; c < 16; c++) h[r * 16 + c] = (uint16_t)r;
  uint16_t *d, *o;
  cudaMalloc(&d, sizeof(h));
  cudaMalloc(&o, 2048);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int box1 : {8, 16})
    for (int start : {0, 3, -1}) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {16, 32};
      cuuint64_t strides[1] = {32};
      cuuint32_t box[2] = {16, (cuuint32_t)box1};
      cuuint32_t es[2] = {1, 2};
      CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (uint32_t tx : {(uint32_t)(box1 * 32), (uint32_t)(box1 * 16)}) {
        cudaMemset(o, 0xff, 2048);
        probe<<<1, 64, 4096>>>(tm, o, start, tx);
        cudaError_t e = cudaDeviceSynchronize();
        uint16_t g[1024];
        cudaMemcpy(g, o, 2048, cudaMemcpyDeviceToHost);
        printf("box1 %2d start %2d expect_tx %3u: encode %d, %s, barrier completed %d; rows in smem (one per 32 B):", box1,
               start, tx, (int)r, cudaGetErrorString(e), (int)g[600]);
        for (int c = 0; c < 20; c++) printf(" %d", (int)g[c * 16]);
        printf("\n");
      }
    }
  return 0;
}
