// Microbenchmark (not part of the product): cycles per tcgen05.mma (cta_group::1, M=128, K=16, fp16->fp32) as a
// function of N, number of interleaved accumulators and the A-descriptor group stride.  Operands are whatever is in
// shared memory (garbage is fine: timing only).  nvcc -gencode arch=compute_100a,code=sm_100a -o mma_bench mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) bench(int N, int nacc, uint32_t sbo_a, int iters, long long *out, int with_stores) {
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
  // zero smem so that the MMAs do not produce NaN/denormal slow paths (there are none, but keep it clean)
  for (uint32_t i = threadIdx.x * 16; i < 200 * 1024; i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s0 + i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t hi_b = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t a0 = s0, b0 = s0 + 96 * 1024;
    // descriptors precomputed; 16 MMAs issued straight-line per loop trip (4 K-steps x 4 operand stages)
    uint64_t ad[16], bd[16];
    uint32_t dd[16];
    for (int j = 0; j < 16; j++) {
      uint32_t alo = (((a0 + (j >> 2) * 16384) >> 4) & 0x3FFFu) | (1u << 16);
      uint32_t blo = (((b0 + ((j >> 2) & 1) * 32768) >> 4) & 0x3FFFu) | (1u << 16);
      ad[j] = ((uint64_t)hi_a << 32) | (alo + 2 * (j & 3));
      bd[j] = ((uint64_t)hi_b << 32) | (blo + 2 * (j & 3));
      dd[j] = tmem + (uint32_t)(j % nacc) * (512 / nacc);
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
  } else if (with_stores && threadIdx.x >= 32) {
    // competing shared-memory writers (emulates TMA fill traffic): ~with_stores x 16 B per thread per loop
    for (int it = 0; it < iters * with_stores; it++)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s0 + 160 * 1024 + ((threadIdx.x * 16 + it * 2048) & 0x7FFF)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 500;
  int Ns[] = {32, 48, 96, 128, 192, 256};
  for (int grid : {1, 148})
    for (int N : Ns)
      for (int nacc : {1, 2})
        for (uint32_t sbo : {1024u, 2048u}) {
          if (nacc == 2 && N > 256) continue;
          for (int ws : {0, 8}) {
            bench<<<grid, 128, 205 * 1024>>>(N, nacc, sbo, iters, d, ws);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
            printf("grid %3d N %3d nacc %d sboA %4u stores %d : %7.1f cycles/MMA (floor N/2 = %d)  %s\n", grid, N, nacc, sbo, ws,
                   (double)mx / (iters * 4), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
        }
  return 0;
}
