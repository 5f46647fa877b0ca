// Microbenchmark (not part of the product): cycles per tcgen05.mma (cta_group::1, M=128, K=16, fp16->fp32) for the
// three K-major swizzle layouts (128B / 64B / 32B rows) as a function of N and of the A-descriptor stride between
// 8-row groups (SBO), with and without the per-tap start shifts of the halo convolution.  Timing only: operands are
// zeros.  nvcc -gencode arch=compute_100a,code=sm_100a -o mma_layout_bench mma_layout_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// row_bytes: 128 / 64 / 32 (layout type 2 / 4 / 6).  sbo_a: bytes between 8-row groups of A.  shift: 1 = the 16
// unrolled MMAs start at pixel offsets (j % 3) rows, as the taps of the halo convolution do.
__global__ void __launch_bounds__(128, 1) bench(int N, int row_bytes, uint32_t sbo_a, int shift, int iters, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x * 16; i < 200 * 1024; i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s0 + i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0) {
    const uint32_t lt = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (lt << 29);
    const uint32_t hi_b = (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (lt << 29);
    const uint32_t a0 = s0, b0 = s0 + 128 * 1024;
    const int ksteps = row_bytes / 32;
    uint64_t ad[16], bd[16];
    uint32_t dd[16];
    for (int j = 0; j < 16; j++) {
      const int k = j % ksteps, blk = j / ksteps;
      uint32_t alo = (((a0 + (blk & 3) * 24576 + (shift ? (blk % 3) * row_bytes : 0)) >> 4) & 0x3FFFu) | (1u << 16);
      uint32_t blo = (((b0 + (blk & 1) * 32768) >> 4) & 0x3FFFu) | (1u << 16);
      ad[j] = ((uint64_t)hi_a << 32) | (alo + 2 * k);
      bd[j] = ((uint64_t)hi_b << 32) | (blo + 2 * k);
      dd[j] = tmem + (uint32_t)(j & 1) * 256;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters / 4; it++) {
#pragma unroll
      for (int j = 0; j < 16; j++) mma(dd[j], ad[j], bd[j], idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 400;
  for (int row_bytes : {128, 64, 32})
    for (int N : {32, 48, 96})
      for (int mult : {8, 10, 11, 12, 16, 17, 18, 20})   // SBO of A in rows (pixels): 8 = canonical, 16 = halo tile of 16 pixels
        for (int shift : {0, 1}) {
          const uint32_t sbo = (uint32_t)mult * row_bytes;
          if (sbo % 16) continue;
          bench<<<148, 128, 205 * 1024>>>(N, row_bytes, sbo, shift, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[148];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
          printf("row %3d B  N %3d  sboA %2d rows (%4u B) shift %d : %6.1f cycles/MMA  %s\n", row_bytes, N, mult, sbo, shift,
                 (double)mx / (iters * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
