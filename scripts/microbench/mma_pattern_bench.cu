// Microbenchmark (not part of the product): what slows short accumulation chains?  M128 x N x K16 MMAs are issued in
// groups of G into a ring of accumulators (first MMA of a group overwrites), every group followed by a
// tcgen05.commit, optionally with `ldwarps` other warps reading TMEM with tcgen05.ld all the time (the epilogue's
// traffic).  nvcc -gencode arch=compute_100a,code=sm_100a -o mma_pattern_bench mma_pattern_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int G>
__global__ void __launch_bounds__(512, 1) bench(int N, int nacc, int acc_stride, int commits, int ldwarps, int groups,
                                                long long *out, const uint8_t *gsrc, int bulk_kb) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x * 16; i < 128 * 1024; i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s0 + i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = ((256u >> 4) & 0x3FFFu) | (1u << 14) | (6u << 29);   // 32-byte rows (a0's layout)
    uint64_t ad[G], bd[G];
    for (int j = 0; j < G; j++) {
      ad[j] = ((uint64_t)hi << 32) | ((((s0 + (j % 3) * 32 + (j / 3) * 512) >> 4) & 0x3FFFu) | (1u << 16));
      bd[j] = ((uint64_t)hi << 32) | ((((s0 + 64 * 1024 + j * 2048) >> 4) & 0x3FFFu) | (1u << 16));
    }
    long long t0 = clock64();
    int a = 0;
    for (int g = 0; g < groups; g++) {
      const uint32_t d = tmem + (uint32_t)a * acc_stride;
#pragma unroll
      for (int j = 0; j < G; j++) mma(d, ad[j], bd[j], idesc, j != 0);
      if (commits)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      if (commits > 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      if (++a == nacc) a = 0;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (warp == 1 && bulk_kb > 0 && (threadIdx.x & 31) == 0) {
    // TMA-like fill traffic: bulk_kb KB per round into the upper part of shared memory, back to back until stopped
    const uint32_t b3 = smem_u32(&bar3);
    uint32_t phase = 0;
    const uint8_t *src = gsrc + (size_t)blockIdx.x * 64 * 1024;
    while (!stop) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b3), "r"((uint32_t)bulk_kb * 1024u) : "memory");
      for (int k = 0; k < bulk_kb; k += 4)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s0 + 96u * 1024u + (uint32_t)k * 1024u), "l"(src + (size_t)k * 1024), "r"(4096u), "r"(b3) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b3), "r"(phase) : "memory");
      phase ^= 1u;
    }
  } else if (warp >= 4 && warp < 4 + ldwarps) {
    // epilogue-like TMEM readers: 8 columns at a time over the accumulator ring, lanes of this warp's quadrant
    const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    int col = 0;
    while (!stop) {
      uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(row + col));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
      col = (col + 8) & 255;
    }
    if (acc == 0x12345) out[147] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long *d;
  uint8_t *g;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMalloc(&g, (size_t)148 * 64 * 1024);
  cudaMemset(g, 0, (size_t)148 * 64 * 1024);
  cudaFuncSetAttribute(bench<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  cudaFuncSetAttribute(bench<36>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  const int groups = 200;
  for (int N : {32, 48})
    for (int commits : {0, 2})
      for (int ldwarps : {0, 12})
        for (int bulk_kb : {0, 8, 32})
          for (int G : {9, 36}) {
            if (G == 9) bench<9><<<148, 512, 136 * 1024>>>(N, 6, 64, commits, ldwarps, groups, d, g, bulk_kb);
            else bench<36><<<148, 512, 136 * 1024>>>(N, 6, 64, commits, ldwarps, groups, d, g, bulk_kb);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 146; i++) mx = h[i] > mx ? h[i] : mx;
            printf("N %2d G %2d commits %d ldwarps %2d bulk %2d KB/round : %6.1f cycles/MMA  %s\n", N, G, commits, ldwarps,
                   bulk_kb, (double)mx / (groups * G), e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
  return 0;
}
