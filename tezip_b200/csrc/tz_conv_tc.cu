// tezip_b200 -- PredNet convolutions as implicit GEMMs on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves every 3x3 convolution of `next()` (prednet.py:255-258,290):
//   M = 128 output pixels (a TB x TH x TW patch of the NHWC activation tensor), N = up to 256 output channels,
//   K = 9 taps x Cin, walked tap-major in blocks of KC channels.
//   warp 0  : TMA producer.  The A operand of tap (dy,dx) is the SAME 4-D box of the fp16 activation tensor
//             shifted by (dy-1, dx-1); TMA zero-fills out-of-bounds rows/columns, which is exactly Keras'
//             'same' padding, so no im2col buffer exists anywhere.  The B operand is a [N_tile x KC] box of the
//             pre-packed K-major weight matrix.  Both land in 128B/64B/32B-swizzled shared memory.
//   warp 1  : allocates TMEM and issues tcgen05.mma (M128 x N x K16, fp16 x fp16 -> fp32 in TMEM) from one
//             elected lane; tcgen05.commit releases smem stages / publishes the accumulator.
//   warps 4..: epilogue (4, 8 or 12 warps; warp w reads TMEM lane quadrant w % 4 and every (EW/4)-th 8-channel
//             chunk).  tcgen05.ld the accumulator rows (one pixel per thread) and finish the layer in
//             registers: bias + relu + 2x2 max-pool + error units (A path, prednet.py:274-277,290-291), or
//             bias-map + hard-sigmoid/tanh LSTM cell (R path, prednet.py:255-259) with the nearest-neighbour
//             up-sampling (prednet.py:263-264) folded into the store.  Two TMEM accumulator stages overlap the
//             epilogue of tile i with the MMAs of tile i+1.
// Determinism: one kernel configuration per layer, no split-K, no atomics; each output element is a fixed-order
// sum over K inside the tensor core, independent of batch size, tile position and grid size.
#include "tz_prednet.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#ifndef TZ_EPI_DEBUG
#define TZ_EPI_DEBUG 0
#endif

namespace tz {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The only arrive issued by a thread is "this accumulator has been read out of TMEM" (epilogue -> MMA issuer).  No
// generic-proxy data travels with it -- tcgen05.wait::ld has completed the reads, tcgen05.fence::before_thread_sync
// orders them -- so it is RELAXED: the default .release compiles to MEMBAR.ALL.CTA (.cluster: MEMBAR.ALL.GPU), which
// made every epilogue warp wait, once per tile, until its up-sampled global stores were visible GPU-wide (16 % of
// the LSTM epilogue's stall samples on gates1).
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Wait for the phase with the given parity.  A pipeline bug must surface as an error, never as a hung GPU:
// after ~4 s of waiting the kernel traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  if (ok) return;   // fast path: already complete
  uint64_t t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// one lane of a converged warp (elect.sync): the canonical way to issue single-thread tcgen05 / TMA instructions from
// warp-uniform code -- descriptors computed by all lanes stay in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 8 consecutive fp32 columns: thread i of the warp receives TMEM lane (base lane + i)
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
// 256-bit read-only global load (sm_100): one request per thread instead of two 128-bit ones -- the epilogue reads
// one private 128-byte line per thread and chunk, so its cost is the number of load wavefronts, not the bytes.
__device__ __forceinline__ void ldg256(const float *p, float v[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) helpers: the two CTAs of a cluster run ONE M256 MMA per instruction; each holds its own
// 128 rows of A and half of the B tile, the leader (rank 0) issues, barriers are signalled across the pair.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta_rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_bar), "r"(cta_rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-parity bit: the address then names the leader's smem
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc2_commit(uint32_t bar) {   // arrives on the same barrier offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

// ------------------------------------------------------------------------------------------------ kernel arguments
constexpr int TC_MAX_THREADS = 512;   // 4 role warps (producer, MMA, 2 idle) + up to 12 epilogue warps
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_MAX_ACC = 8;       // TMEM accumulator stages (2 for wide tiles, up to 8 for narrow ones)
constexpr int TC_TMEM_COLS = 512;

struct ConvArgs {
  int B, H, W;
  int tw_log, th_log, tb_log;    // tile = 2^tb x 2^th x 2^tw pixels = 128
  int tiles_w, tiles_h, n_tiles_n;
  uint32_t m_tiles_w, m_tiles_h, m_tiles_n;   // fast_div magic numbers of the three
  int kchunks, ksteps;           // K blocks per tap, MMAs (K=16) per K block
  int ksteps_last;               // pair mode: K=16 steps of the LAST block that hold a real input channel (the rest
                                 // multiply zero weights and are not issued; a1 of the (3,48,96,192) net: 96 = 64 + 32)
  int KC, cin_pad;               // channels per K block, kchunks*KC
  int n_tile;                    // MMA N (multiple of 16)
  int stages;
  int epi_warps;                 // 4, 8 or 12
  int epi_groups;                // 1: every epilogue warp works on every tile (8-channel chunks are shared out);
                                 // G > 1: narrow tiles -- warp group g (4 warps) owns the tiles with index % G == g,
                                 // so G epilogues are in flight and their latency overlaps
  int acc_stages, acc_stride;    // TMEM accumulator ring: stages, columns per stage
  int mma_issuers;               // halo + stationary weights: 2 warps issue the MMAs of alternate tiles, else 1
  int halo;                      // 0: im2col; 1: halo + stationary weights; 4: CTA pair + halo (see conv_tc_kernel)
  int tile_h;                    // tile height in pixels
  uint32_t a_slot, a_tx;         // pair mode: bytes of one halo slot (1024-aligned) and of the TMA box
  uint32_t b_block, b_region;    // halo mode: bytes of one [n_tile x 64] weight block; bytes of all 9*kchunks blocks
  uint32_t a_stride, stage_stride, tx_bytes;
  uint32_t desc_hi;              // upper half of the smem matrix descriptor (SBO, version, swizzle)
  uint32_t idesc;
  // --- A epilogue (pool + error units)
  const float *bias;             // [N]
  const float *ahat_next;        // [H/2, W/2, S_next]
  __half *xe_out;                // X_{l+1}: [B, H/2, W/2, xe_cstride], e written at channels [0, 2*S_next)
  int xe_cstride, S_next;
  int perm16;                    // accumulator columns are permuted inside every group of 16 output channels so that a
                                 // lane of the pooling epilogue ends with FOUR consecutive channels (8-byte stores)
  // --- R epilogue (LSTM)
  const float *bm;               // [H, W, 4R];  when bm_packed: [H, W, R/8, 4 gates, 8] (one 128-byte line per
                                 // thread and 8-channel chunk -> eight 16-byte loads of one cache line)
  int bm_packed;
  const float *c0;               // [H, W, R]
  int R, NC, NCp;                // channels, channels per N tile, gate-block column pitch (NC rounded up to 8)
  __half *xr_out;                // X_{l-1}: [B, 2H, 2W, xr_cstride], r written up-sampled at channel xr_coff
  int xr_cstride, xr_coff;
  float *r0_out;                 // layer 0: r as fp32 [B, H, W, R] (xr_out == nullptr)
  int xr_up;                     // 1: r is written 2x up-sampled (four stores); 0: once, at its own resolution
  // --- raw epilogue (EPI 2): the accumulator row as fp32 [B, H, W, g_cols] (G of the layer-0 split, tc_create)
  float *g_out;
  int g_cols;
  long long *dbg;                // optional [gridDim][8]: MMA-thread cycle breakdown (TZ_CONV_DEBUG), else nullptr
};

__device__ __forceinline__ float hsig(float x) {
  return fminf(fmaxf(__fadd_rn(__fmul_rn(0.2f, x), 0.5f), 0.0f), 1.0f);
}
// tanh(x) = 1 - 2/(exp(2x)+1) on the SFU (ex2.approx + rcp.approx): absolute error ~1e-7, far below the fp16
// operand rounding of the convolutions, saturates correctly to +-1, and is deterministic on a given GPU.
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// Halo mode: the 9 x KSTEPS MMAs of one channel chunk, fully unrolled.  One thread issues every tcgen05.mma of the
// CTA, so for narrow tiles (N <= 64, where an MMA lasts only ~30-100 cycles) the issue loop itself is the
// critical path: descriptors are formed with one 32-bit add each, no loop-carried integer division.
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// n / d for small operands without the ~20-instruction runtime division: m = ceil(2^32 / d) (host), exact while
// n * d < 2^32; d == 1 is encoded as m == 0.
__device__ __forceinline__ int fast_div(int n, uint32_t m) { return m ? (int)__umulhi((uint32_t)n, m) : n; }

template <int KSTEPS>
__device__ __forceinline__ void issue_halo_chunk(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t row16 /*row bytes >> 4*/,
                                                 uint32_t b_tap_stride16, uint32_t idesc, bool first_chunk) {
#pragma unroll
  for (int tap = 0; tap < 9; tap++) {
    const uint32_t al = a_lo + (uint32_t)((tap / 3) * 16 + (tap % 3)) * row16;
    const uint32_t bl = b_lo + (uint32_t)tap * b_tap_stride16;
#pragma unroll
    for (int k = 0; k < KSTEPS; k++)
      tc_mma_f16(d_tmem, make_desc(al + 2 * k, a_hi), make_desc(bl + 2 * k, b_hi), idesc,
                 (tap | k) != 0 || !first_chunk);
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
// TWO = CTA-pair instantiation (launched as clusters of 2): kernels that contain cta_group::2 instructions can only
// be launched with a matching cluster size, so the one-CTA modes use the TWO = false instantiation.
template <int EPI, bool TWO>  // EPI 0: A path (pool + E), 1: R path (LSTM)
__global__ void __launch_bounds__(TC_MAX_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvArgs P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 2 * TC_MAX_ACC + 1 + 4];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full0 = smem_u32(&bars[0]);
  const uint32_t empty0 = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const uint32_t tempty0 = smem_u32(&bars[2 * TC_MAX_STAGES + TC_MAX_ACC]);
  const uint32_t bfull = smem_u32(&bars[2 * TC_MAX_STAGES + 2 * TC_MAX_ACC]);   // halo mode: all weights have landed
  const uint32_t afull0 = smem_u32(&bars[2 * TC_MAX_STAGES + 2 * TC_MAX_ACC + 1]);   // pair mode: 2 halo slots
  const uint32_t aempty0 = afull0 + 16;

  constexpr bool two = TWO;                               // CTA-pair mode (launched as clusters of 2)
  uint32_t crank = 0;                                     // 0 = leader (issues the MMAs)
  if constexpr (TWO) crank = cluster_ctarank();
  const int tiles_img = P.tiles_w * P.tiles_h;
  const int tiles_b = (P.B + (1 << P.tb_log) - 1) >> P.tb_log;
  // scheduled units: (M tile, N tile) per CTA, or (pair of M tiles, N tile) per cluster; unit index steps by the
  // number of CTAs / clusters.  m_of(t) is the M tile this CTA owns inside unit t.
  const int n_tiles = two ? ((tiles_img * tiles_b + 1) >> 1) * P.n_tiles_n : tiles_img * tiles_b * P.n_tiles_n;
  const int t_first = two ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = two ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kblocks = 9 * P.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < P.acc_stages; a++) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, (uint32_t)(two ? 2 * P.epi_warps : (P.epi_groups > 1 ? 4 : P.epi_warps)));
    }
    mbar_init(bfull, 1);
    for (int a = 0; a < 2; a++) {
      mbar_init(afull0 + 8 * a, 1);
      mbar_init(aempty0 + 8 * a, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "r"((uint32_t)TC_TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "r"((uint32_t)TC_TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // Ring state is kept incrementally (stage index + phase bit) and descriptors are formed with 32-bit adds: ONE
  // thread issues every TMA and ONE thread every tcgen05.mma of the CTA, so the instruction count of these two
  // loops bounds the tensor pipe (a microbenchmark reaches N/2 cycles per M128 MMA only with straight-line issue;
  // runtime divisions and 64-bit descriptor arithmetic per MMA cost more than the MMA itself).
  // Halo + stationary-weights mode: the MMA issue loop.  The tensor work of one tile is short here (9 * ksteps MMAs
  // of <= 56 cycles per chunk), shorter than the ~1000 cycles one thread needs to walk the barriers and form the
  // descriptors of a tile (measured: a0 issued at 119 cycles per MMA against a floor of 44), so TWO warps issue,
  // each taking every other tile of this CTA (any thread may issue tcgen05.mma; a tcgen05.commit tracks the MMAs
  // of the thread that executes it, which is exactly the per-tile granularity the barriers need).
  auto halo1_issue = [&](uint32_t k0, uint32_t kstep, bool dbg) {
    if ((int)blockIdx.x >= n_tiles) return;
    const uint32_t hi = P.desc_hi, idesc = P.idesc;
    const uint32_t rowb = 2u * (uint32_t)P.KC;
    const uint32_t ahi_halo = (((16u * rowb) >> 4) & 0x3FFFu) | (hi & 0xFFFFC000u);
    const uint32_t btap16 = ((uint32_t)P.kchunks * P.b_block) >> 4;
    const uint32_t nst = (uint32_t)P.stages, nacc = (uint32_t)P.acc_stages, kch = (uint32_t)P.kchunks;
    auto advance = [](uint32_t &idx, uint32_t &phase, uint32_t by, uint32_t size) {
      idx += by;
      while (idx >= size) {
        idx -= size;
        phase ^= 1u;
      }
    };
    uint32_t s = 0, ph = 0, a = 0, aph = 0;
    advance(s, ph, k0 * kch, nst);
    advance(a, aph, k0, nacc);
    long long w_tempty = 0, w_full = 0, c0 = 0, t_start = 0;
    if (dbg) t_start = clock64();
    mbar_wait(bfull, 0);
    tc_fence_after();
    for (int t = t_first + (int)k0 * t_step; t < n_tiles; t += (int)kstep * t_step) {
      if (dbg) c0 = clock64();
      mbar_wait(tempty0 + 8 * a, aph ^ 1u);
      if (dbg) w_tempty += clock64() - c0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (uint32_t)P.acc_stride;
      for (uint32_t ch = 0; ch < kch; ch++) {
        if (dbg) c0 = clock64();
        mbar_wait(full0 + 8 * s, ph);
        if (dbg) w_full += clock64() - c0;
        tc_fence_after();
        const uint32_t sa = smem0 + P.b_region + s * P.stage_stride;
        const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = (((smem0 + ch * P.b_block) >> 4) & 0x3FFFu) | (1u << 16);
        if (elect_one()) {
          if (P.ksteps == 4)
            issue_halo_chunk<4>(d_tmem, a_lo, ahi_halo, b_lo, hi, rowb >> 4, btap16, idesc, ch == 0);
          else if (P.ksteps == 2)
            issue_halo_chunk<2>(d_tmem, a_lo, ahi_halo, b_lo, hi, rowb >> 4, btap16, idesc, ch == 0);
          else
            issue_halo_chunk<1>(d_tmem, a_lo, ahi_halo, b_lo, hi, rowb >> 4, btap16, idesc, ch == 0);
          tc_commit(empty0 + 8 * s);   // halo slot free once its MMAs retire
          if (ch + 1 == kch) tc_commit(tfull0 + 8 * a);   // accumulator complete
        }
        __syncwarp();
        advance(s, ph, 1, nst);
      }
      advance(s, ph, (kstep - 1) * kch, nst);   // the other issuer's chunks
      advance(a, aph, kstep, nacc);
    }
    if (dbg) {
      long long *o = P.dbg + 8 * blockIdx.x;
      o[0] = clock64() - t_start;
      o[1] = w_tempty;
      o[2] = w_full;
      o[3] = 0;
    }
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    // (whole warp walks the loop with warp-uniform state; one elected lane issues the TMA instructions)
    {
      uint32_t s = 0, ph = 0;      // main ring (im2col: A+B stages; halo: halo slots; pair: weight stages)
      uint32_t sl = 0, pha = 0;    // pair mode: halo slot ring
      if (P.halo == 1 && (int)blockIdx.x < n_tiles) {
        // Halo mode: the whole weight matrix of this CTA's N tile stays resident in shared memory (loaded once)
        const int n0 = ((int)blockIdx.x % P.n_tiles_n) * P.n_tile;
        if (elect_one()) {
          mbar_expect_tx(bfull, (uint32_t)kblocks * (uint32_t)P.n_tile * (uint32_t)(2 * P.KC));
          for (int kb = 0; kb < kblocks; kb++)   // block kb = tap * kchunks + chunk
            tma_load_2d(smem0 + kb * P.b_block, &tmB, bfull, kb * P.KC, n0);
        }
        __syncwarp();
      }
      const uint32_t pair_bbase = smem0 + 2 * P.a_slot;
      for (int t = t_first; t < n_tiles; t += t_step) {
        int mt = fast_div(t, P.m_tiles_n);
        const int nt = t - mt * P.n_tiles_n;
        if (two) mt = 2 * mt + (int)crank;
        const int mq = fast_div(mt, P.m_tiles_w);
        const int twi = mt - mq * P.tiles_w;
        const int tbi = fast_div(mq, P.m_tiles_h);
        const int thi = mq - tbi * P.tiles_h;
        const int w0 = twi << P.tw_log, h0 = thi * P.tile_h, b0 = tbi << P.tb_log;
        const int n0 = nt * P.n_tile;
        if constexpr (TWO) {
          // CTA-pair mode: this CTA fetches the halo box of ITS M tile and ITS half of every weight block; all
          // transaction bytes of the pair are counted on the leader's barriers (the leader alone waits on them).
          int kc = 0;
          const int nb0 = n0 + (int)crank * (P.n_tile >> 1);
          for (int c = 0; c < P.cin_pad; c += P.KC, kc += P.KC) {
            mbar_wait(aempty0 + 8 * sl, pha ^ 1u);
            if (elect_one()) {
              if (crank == 0) mbar_expect_tx(afull0 + 8 * sl, 2 * P.a_tx);
              tma2_load_4d(smem0 + sl * P.a_slot, &tmA, afull0 + 8 * sl, c, w0 - 1, h0 - 1, b0);
            }
            __syncwarp();
            sl ^= 1u;
            pha ^= (sl == 0);
            int kcoord = kc;
            for (int tg = 0; tg < 3; tg++) {   // one weight stage = the three taps of one filter row
              mbar_wait(empty0 + 8 * s, ph ^ 1u);
              if (elect_one()) {
                if (crank == 0) mbar_expect_tx(full0 + 8 * s, 2 * P.tx_bytes);
                const uint32_t sb = pair_bbase + s * P.stage_stride;
#pragma unroll
                for (int j = 0; j < 3; j++)
                  tma2_load_2d(sb + (uint32_t)j * P.b_block, &tmB, full0 + 8 * s, kcoord + j * P.cin_pad, nb0);
              }
              __syncwarp();
              kcoord += 3 * P.cin_pad;
              if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
            }
          }
        } else {
        if (P.halo == 0) {
          // im2col mode: per (tap, chunk) one shifted activation box + one weight box
          int kcoord = 0;
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++)
              for (int c = 0; c < P.cin_pad; c += P.KC) {
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                if (elect_one()) {
                  const uint32_t full = full0 + 8 * s;
                  mbar_expect_tx(full, P.tx_bytes);
                  const uint32_t sa = smem0 + s * P.stage_stride;
                  tma_load_4d(sa, &tmA, full, c, w0 + dx, h0 + dy, b0);
                  tma_load_2d(sa + P.a_stride, &tmB, full, kcoord, n0);
                }
                __syncwarp();
                kcoord += P.KC;
                if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
              }
        } else if (P.halo == 1) {
          // halo mode: per chunk ONE box of (TH+2) x 16 pixels; the nine taps read it through shifted descriptors
          for (int c = 0; c < P.cin_pad; c += P.KC) {
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            if (elect_one()) {
              const uint32_t full = full0 + 8 * s;
              mbar_expect_tx(full, P.tx_bytes);
              tma_load_4d(smem0 + P.b_region + s * P.stage_stride, &tmA, full, c, w0 - 1, h0 - 1, b0);
            }
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
          }
        }
        }   // one-CTA modes
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // The whole warp walks the loop (waits, ring state, descriptors are warp-uniform); one elected lane issues.
    if (P.halo == 1) {
      halo1_issue(0, (uint32_t)P.mma_issuers, P.dbg != nullptr && lane == 0);
    } else if (crank == 0) {
      uint32_t s = 0, ph = 0, sl = 0, pha = 0;
      uint32_t a = 0, aph = 0;   // accumulator ring
      const uint32_t hi = P.desc_hi;
      const uint32_t idesc = P.idesc;
      // halo tiles: 16 pixels per image row -> the 8-row groups (one image row of the 8-wide tile) are 16*rowb apart
      // (pair mode: 10 pixels of 128 B).  Neither needs to be a multiple of the swizzle period: the swizzle is a
      // function of the absolute shared-memory address bits (what TMA wrote), and the descriptor base_offset stays 0
      // -- measured on B200: base_offset = dx gives wrong results, 0 matches the im2col path to rounding.
      const uint32_t ahi_pair = ((1280u >> 4) & 0x3FFFu) | (hi & 0xFFFFC000u);
      const uint32_t pair_bbase = smem0 + 2 * P.a_slot;
      long long w_tempty = 0, w_full = 0, w_afull = 0, c0 = 0, t_start = 0;
      const bool dbg = P.dbg != nullptr && lane == 0;
      if (dbg) t_start = clock64();
      for (int t = t_first; t < n_tiles; t += t_step) {
        if (dbg) c0 = clock64();
        mbar_wait(tempty0 + 8 * a, aph ^ 1u);
        if (dbg) w_tempty += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * (uint32_t)P.acc_stride;
        if constexpr (TWO) {
          // one M256 MMA per instruction: rows 0..127 from this CTA's halo box, rows 128..255 from the peer's (same
          // smem offsets), B rows [0, N/2) from this CTA's stage, [N/2, N) from the peer's
          uint32_t acc = 0;
          for (int ch = 0; ch < P.kchunks; ch++) {
            if (dbg) c0 = clock64();
            mbar_wait(afull0 + 8 * sl, pha);
            if (dbg) w_afull += clock64() - c0;
            tc_fence_after();
            const uint32_t a_base = (((smem0 + sl * P.a_slot) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
            for (int tg = 0; tg < 3; tg++) {   // filter row tg: three taps per weight stage
              if (dbg) c0 = clock64();
              mbar_wait(full0 + 8 * s, ph);
              if (dbg) w_full += clock64() - c0;
              tc_fence_after();
              const uint32_t b_base = (((pair_bbase + s * P.stage_stride) >> 4) & 0x3FFFu) | (1u << 16);
              if (elect_one()) {
                if (ch + 1 == P.kchunks && P.ksteps_last == 2) {   // half-empty last block: two K steps per tap
#pragma unroll
                  for (int j = 0; j < 3; j++) {
                    const uint32_t a_lo = a_base + (uint32_t)(tg * 10 + j) * 8u;
                    const uint32_t b_lo = b_base + (uint32_t)j * (P.b_block >> 4);
#pragma unroll
                    for (int k = 0; k < 2; k++)
                      tc2_mma_f16(d_tmem, make_desc(a_lo + 2 * k, ahi_pair), make_desc(b_lo + 2 * k, hi), idesc,
                                  acc | (uint32_t)(tg | j | k));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 3; j++) {   // 64-channel chunks: 128-byte pixels
                    const uint32_t a_lo = a_base + (uint32_t)(tg * 10 + j) * 8u;
                    const uint32_t b_lo = b_base + (uint32_t)j * (P.b_block >> 4);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                      tc2_mma_f16(d_tmem, make_desc(a_lo + 2 * k, ahi_pair), make_desc(b_lo + 2 * k, hi), idesc,
                                  acc | (uint32_t)(tg | j | k));
                  }
                }
                tc2_commit(empty0 + 8 * s);
              }
              __syncwarp();
              if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
            }
            acc = 1;
            if (elect_one()) tc2_commit(aempty0 + 8 * sl);
            __syncwarp();
            sl ^= 1u;
            pha ^= (sl == 0);
          }
          if (elect_one()) tc2_commit(tfull0 + 8 * a);
          __syncwarp();
          if (++a == (uint32_t)P.acc_stages) { a = 0; aph ^= 1u; }
          continue;
        }
        if (P.halo == 0) {
          uint32_t acc = 0;
          for (int kb = 0; kb < kblocks; kb++) {
            if (dbg) c0 = clock64();
            mbar_wait(full0 + 8 * s, ph);
            if (dbg) w_full += clock64() - c0;
            tc_fence_after();
            const uint32_t sa = smem0 + s * P.stage_stride;
            const uint32_t alo = ((sa >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t blo = (((sa + P.a_stride) >> 4) & 0x3FFFu) | (1u << 16);
            // advance 32 bytes (16 fp16 of K) inside the swizzled row per MMA
            if (elect_one()) {
              if (P.ksteps == 4) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                  tc_mma_f16(d_tmem, make_desc(alo + 2 * k, hi), make_desc(blo + 2 * k, hi), idesc, acc | (uint32_t)k);
              } else {
                for (int k = 0; k < P.ksteps; k++)
                  tc_mma_f16(d_tmem, make_desc(alo + 2 * k, hi), make_desc(blo + 2 * k, hi), idesc, acc | (uint32_t)k);
              }
              tc_commit(empty0 + 8 * s);  // frees the smem stage when these MMAs retire
            }
            __syncwarp();
            acc = 1;
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
          }
        }
        if (elect_one()) tc_commit(tfull0 + 8 * a);    // accumulator complete
        __syncwarp();
        if (++a == (uint32_t)P.acc_stages) { a = 0; aph ^= 1u; }
      }
      if (dbg) {
        long long *o = P.dbg + 8 * blockIdx.x;
        o[0] = clock64() - t_start;
        o[1] = w_tempty;
        o[2] = w_full;
        o[3] = w_afull;
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================================================================== second MMA issuer (halo mode only)
    if (P.halo == 1 && P.mma_issuers == 2) halo1_issue(1, 2, false);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================================== epilogue (warps 4 .. 4+epi_warps-1)
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int group = (warp - 4) >> 2;
    const int part = P.epi_groups > 1 ? 0 : group;          // which share of the 8-channel chunks
    const int nparts = P.epi_groups > 1 ? 1 : (P.epi_warps >> 2);
    const int m = quad * 32 + lane;            // accumulator row == pixel inside the tile
    const int tw = m & ((1 << P.tw_log) - 1);
    const int th = (m >> P.tw_log) & ((1 << P.th_log) - 1);
    const int tb = m >> (P.tw_log + P.th_log);
    uint32_t tc = 0;
    uint32_t ring_a = 0, ring_ph = 0, ring_g = 0;   // accumulator stage / phase / owning group of tile tc, kept incrementally
#if TZ_EPI_DEBUG   // compile-time: the LSTM epilogue has no registers to spare (nvcc -DTZ_EPI_DEBUG=1 via TZ_NVCC_FLAGS)
    const bool edbg = P.dbg != nullptr && warp == 4 && lane == 0;
#else
    constexpr bool edbg = false;
#endif
    long long e_wait = 0, e_work = 0, e_tiles = 0, ec0 = 0, e_a = 0, e_b = 0;
    for (int t = t_first; t < n_tiles; t += t_step, tc++) {
      const uint32_t a = ring_a, aph = ring_ph;
      const bool mine = P.epi_groups <= 1 || (int)ring_g == group;
      if (++ring_a == (uint32_t)P.acc_stages) {
        ring_a = 0;
        ring_ph ^= 1u;
      }
      if (++ring_g >= (uint32_t)P.epi_groups) ring_g = 0;
      if (!mine) continue;   // another group's tile
      int mt = fast_div(t, P.m_tiles_n);
      const int nt = t - mt * P.n_tiles_n;
      if (two) mt = 2 * mt + (int)crank;
      const int mq = fast_div(mt, P.m_tiles_w);
      const int twi = mt - mq * P.tiles_w;
      const int tbi = fast_div(mq, P.m_tiles_h);
      const int thi = mq - tbi * P.tiles_h;
      const int w_tile = (twi << P.tw_log) + tw, b = (tbi << P.tb_log) + tb;
      if (edbg) ec0 = clock64();
      mbar_wait(tfull0 + 8 * a, aph);
      if (edbg) {
        const long long now = clock64();
        e_wait += now - ec0;
        ec0 = now;
        e_tiles++;
      }
      tc_fence_after();
      const int h = thi * P.tile_h + th;
      const int w = w_tile;
      const bool valid = (b < P.B) && (h < P.H) && (w < P.W);
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + a * (uint32_t)P.acc_stride;
      if (EPI == 0) {
        // a = maxpool2x2(relu(conv + bias));  e = [relu(ahat - a), relu(a - ahat)]  -> fp16 into X_{l+1}
        const int Ho = P.H >> 1, Wo = P.W >> 1;
        const bool writer = valid && ((tw & 1) == 0) && ((th & 1) == 0);
        const long long opix = ((long long)b * Ho + (h >> 1)) * Wo + (w >> 1);
        const float *ah = P.ahat_next + ((long long)(h >> 1) * Wo + (w >> 1)) * P.S_next;
        __half *dst = P.xe_out + opix * P.xe_cstride;
        const int n_real = P.S_next - nt * P.n_tile < P.n_tile ? P.S_next - nt * P.n_tile : P.n_tile;
        constexpr int PF = 6;
        const bool vec_ok = ((P.S_next | P.xe_cstride) & 7) == 0;
        if (P.perm16) {
          // Lane-balanced pooling over PAIRS of 8-channel chunks (columns [16k, 16k + 16)).  As below, the four lanes
          // of a 2x2 window swap halves (xor 1 keeps 4 of a chunk's 8 columns, xor tile-width keeps 2), so a lane ends
          // with the pooled maximum of columns (co, co + 1) of both chunks -- and the weight rows were packed so that
          // these four columns are the four CONSECUTIVE channels 16k + 2co .. + 3 (make_conv): bias and A-hat arrive
          // as one 16-byte load each, E+ and E- leave as one 8-byte store each.  The epilogue of the narrow layer-0
          // convolution is bound by the L1 data pipe (shuffles + sector-granular stores), not by warps or issue
          // slots; this halves its store wavefronts.
          const int hm = 1 << P.tw_log;
          const bool wbit = (lane & 1) != 0, hbit = (lane & hm) != 0;
          const int co = (wbit ? 4 : 0) + (hbit ? 2 : 0);
          const int npairs = n_real >> 4;
          constexpr int PP = PF / 2;
          float4 ahq[PP];
#pragma unroll
          for (int c = 0; c < PP; c++) {
            const int k = part + c * nparts;
            if (valid && k < npairs) ahq[c] = __ldg(reinterpret_cast<const float4 *>(ah + nt * P.n_tile + 16 * k + 2 * co));
          }
          auto finish_pair = [&](int k, const float *va, const float *vb, float4 aa) {
            const int ch4 = nt * P.n_tile + 16 * k + 2 * co;
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(P.bias + ch4));
            float ua[4], ub[4], z[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const float sa = wbit ? va[q] : va[q + 4], ka = wbit ? va[q + 4] : va[q];
              const float sb = wbit ? vb[q] : vb[q + 4], kb = wbit ? vb[q + 4] : vb[q];
              ua[q] = fmaxf(ka, __shfl_xor_sync(0xffffffffu, sa, 1));
              ub[q] = fmaxf(kb, __shfl_xor_sync(0xffffffffu, sb, 1));
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
              const float sa = hbit ? ua[q] : ua[q + 2], ka = hbit ? ua[q + 2] : ua[q];
              const float sb = hbit ? ub[q] : ub[q + 2], kb = hbit ? ub[q + 2] : ub[q];
              z[q] = fmaxf(ka, __shfl_xor_sync(0xffffffffu, sa, hm));
              z[2 + q] = fmaxf(kb, __shfl_xor_sync(0xffffffffu, sb, hm));
            }
            if (valid) {
              const float a0 = fmaxf(__fadd_rn(z[0], bb.x), 0.0f), a1 = fmaxf(__fadd_rn(z[1], bb.y), 0.0f);
              const float a2 = fmaxf(__fadd_rn(z[2], bb.z), 0.0f), a3 = fmaxf(__fadd_rn(z[3], bb.w), 0.0f);
              const __half2 p0 = __floats2half2_rn(fmaxf(__fsub_rn(aa.x, a0), 0.0f), fmaxf(__fsub_rn(aa.y, a1), 0.0f));
              const __half2 p1 = __floats2half2_rn(fmaxf(__fsub_rn(aa.z, a2), 0.0f), fmaxf(__fsub_rn(aa.w, a3), 0.0f));
              const __half2 m0 = __floats2half2_rn(fmaxf(__fsub_rn(a0, aa.x), 0.0f), fmaxf(__fsub_rn(a1, aa.y), 0.0f));
              const __half2 m1 = __floats2half2_rn(fmaxf(__fsub_rn(a2, aa.z), 0.0f), fmaxf(__fsub_rn(a3, aa.w), 0.0f));
              uint2 up, dn;
              up.x = *reinterpret_cast<const uint32_t *>(&p0);
              up.y = *reinterpret_cast<const uint32_t *>(&p1);
              dn.x = *reinterpret_cast<const uint32_t *>(&m0);
              dn.y = *reinterpret_cast<const uint32_t *>(&m1);
              *reinterpret_cast<uint2 *>(dst + ch4) = up;
              *reinterpret_cast<uint2 *>(dst + P.S_next + ch4) = dn;
            }
          };
#pragma unroll
          for (int c = 0; c < PP; c++) {
            const int k = part + c * nparts;
            if (k < npairs) {   // warp-uniform
              float va[8], vb[8];
              tc_ld8(trow + 16 * k, va);
              tc_ld8(trow + 16 * k + 8, vb);
              tc_ld_wait();
              finish_pair(k, va, vb, ahq[c]);
            }
          }
          for (int k = part + PP * nparts; k < npairs; k += nparts) {   // more than PP pairs per warp
            float4 a4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (valid) a4 = __ldg(reinterpret_cast<const float4 *>(ah + nt * P.n_tile + 16 * k + 2 * co));
            float va[8], vb[8];
            tc_ld8(trow + 16 * k, va);
            tc_ld8(trow + 16 * k + 8, vb);
            tc_ld_wait();
            finish_pair(k, va, vb, a4);
          }
        } else
        if (vec_ok && ((P.H | P.W) & 1) == 0 && P.tw_log >= 1 && P.th_log >= 1) {
          // Lane-balanced pooling.  The four lanes of a 2x2 window swap halves instead of all computing everything
          // (transpose-reduce: xor 1 keeps 4 of the 8 channels, xor tile-width keeps 2), so each lane ends with the
          // pooled maximum of TWO channels and does bias, relu, the error units and a 4-byte store for those: 6
          // shuffles per chunk instead of 16, no idle lanes in the E math.  max and (+bias, relu) commute exactly
          // (both monotone), so the result is bit-identical to pooling relu(conv + bias).
          const int hm = 1 << P.tw_log;
          const bool wbit = (lane & 1) != 0, hbit = (lane & hm) != 0;
          const int co = (wbit ? 4 : 0) + (hbit ? 2 : 0);   // this lane's channel pair inside an 8-channel chunk
          float2 ahq[PF];
#pragma unroll
          for (int c = 0; c < PF; c++) {   // Ahat0 of all chunks up front: the L2 latency is paid once per tile
            const int j0 = 8 * (part + c * nparts);
            if (valid && j0 < n_real) ahq[c] = __ldg(reinterpret_cast<const float2 *>(ah + nt * P.n_tile + j0 + co));
          }
          // Two chunks at a time: their shuffle / max / store chains are independent, which gives the scheduler two
          // instruction streams per warp (the epilogue runs 3 warps per scheduler; with one chain per warp it was
          // bound by dependent-issue latency, ~700-1100 cycles per chunk).
          auto finish2 = [&](int ja, int jb, bool has_b, const float *va, const float *vb, float2 aa, float2 ab) {
            const int cha = nt * P.n_tile + ja + co, chb = nt * P.n_tile + jb + co;
            const float2 ba = __ldg(reinterpret_cast<const float2 *>(P.bias + cha));
            const float2 bb = has_b ? __ldg(reinterpret_cast<const float2 *>(P.bias + chb)) : make_float2(0.0f, 0.0f);
            float ua[4], ub[4], za[2], zb[2];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float sa = wbit ? va[k] : va[k + 4], ka = wbit ? va[k + 4] : va[k];
              const float sb = wbit ? vb[k] : vb[k + 4], kb = wbit ? vb[k + 4] : vb[k];
              ua[k] = fmaxf(ka, __shfl_xor_sync(0xffffffffu, sa, 1));
              ub[k] = fmaxf(kb, __shfl_xor_sync(0xffffffffu, sb, 1));
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const float sa = hbit ? ua[k] : ua[k + 2], ka = hbit ? ua[k + 2] : ua[k];
              const float sb = hbit ? ub[k] : ub[k + 2], kb = hbit ? ub[k + 2] : ub[k];
              za[k] = fmaxf(ka, __shfl_xor_sync(0xffffffffu, sa, hm));
              zb[k] = fmaxf(kb, __shfl_xor_sync(0xffffffffu, sb, hm));
            }
            if (valid) {
              const float a0 = fmaxf(__fadd_rn(za[0], ba.x), 0.0f), a1 = fmaxf(__fadd_rn(za[1], ba.y), 0.0f);
              const float b0 = fmaxf(__fadd_rn(zb[0], bb.x), 0.0f), b1 = fmaxf(__fadd_rn(zb[1], bb.y), 0.0f);
              *reinterpret_cast<__half2 *>(dst + cha) =
                  __floats2half2_rn(fmaxf(__fsub_rn(aa.x, a0), 0.0f), fmaxf(__fsub_rn(aa.y, a1), 0.0f));
              *reinterpret_cast<__half2 *>(dst + P.S_next + cha) =
                  __floats2half2_rn(fmaxf(__fsub_rn(a0, aa.x), 0.0f), fmaxf(__fsub_rn(a1, aa.y), 0.0f));
              if (has_b) {
                *reinterpret_cast<__half2 *>(dst + chb) =
                    __floats2half2_rn(fmaxf(__fsub_rn(ab.x, b0), 0.0f), fmaxf(__fsub_rn(ab.y, b1), 0.0f));
                *reinterpret_cast<__half2 *>(dst + P.S_next + chb) =
                    __floats2half2_rn(fmaxf(__fsub_rn(b0, ab.x), 0.0f), fmaxf(__fsub_rn(b1, ab.y), 0.0f));
              }
            }
          };
#pragma unroll
          for (int c = 0; c < PF; c += 2) {
            const int ja = 8 * (part + c * nparts), jb = ja + 8 * nparts;
            if (ja < n_real) {   // warp-uniform
              const bool has_b = jb < n_real;
              float va[8], vb[8];
              tc_ld8(trow + ja, va);
              tc_ld8(trow + (has_b ? jb : ja), vb);
              tc_ld_wait();
              finish2(ja, jb, has_b, va, vb, ahq[c], ahq[c + 1]);
            }
          }
          for (int j0 = 8 * (part + PF * nparts); j0 < n_real; j0 += 8 * nparts) {   // more than PF chunks per warp
            float2 a2 = make_float2(0.0f, 0.0f);
            if (valid) a2 = __ldg(reinterpret_cast<const float2 *>(ah + nt * P.n_tile + j0 + co));
            float va[8];
            tc_ld8(trow + j0, va);
            tc_ld_wait();
            finish2(j0, j0, false, va, va, a2, a2);
          }
        } else {
        // Generic path (odd sizes, unaligned channel counts): every lane pools all 8 channels, the lane of the
        // window's top-left pixel writes.  Ahat0 of all of this warp's chunks is requested up front.
        float ahp[PF][8];
#pragma unroll
        for (int c = 0; c < PF; c++) {
          const int j0 = 8 * (part + c * nparts);
          const int ch0 = nt * P.n_tile + j0;
          if (writer && vec_ok && j0 < n_real && ch0 + 8 <= P.S_next) ldg256(ah + ch0, ahp[c]);
        }
        auto chunk = [&](int j0, const float *ahv) {
          float v[8];
          tc_ld8(trow + j0, v);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const int ch = nt * P.n_tile + j0 + j;
            float x = (ch < P.S_next) ? fmaxf(__fadd_rn(v[j], P.bias[ch]), 0.0f) : 0.0f;
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1 << P.tw_log));
            v[j] = x;
          }
          if (writer) {
            const int ch0 = nt * P.n_tile + j0;
            if (ch0 + 8 <= P.S_next && vec_ok) {
              float al[8];
              if (!ahv) ldg256(ah + ch0, al);
              __align__(16) __half up[8], dn[8];
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const float a_ = ahv ? ahv[j] : al[j];
                up[j] = __float2half_rn(fmaxf(__fsub_rn(a_, v[j]), 0.0f));
                dn[j] = __float2half_rn(fmaxf(__fsub_rn(v[j], a_), 0.0f));
              }
              *reinterpret_cast<uint4 *>(dst + ch0) = *reinterpret_cast<const uint4 *>(up);
              *reinterpret_cast<uint4 *>(dst + P.S_next + ch0) = *reinterpret_cast<const uint4 *>(dn);
            } else {
              for (int j = 0; j < 8 && ch0 + j < P.S_next; j++) {
                const float ahj = ah[ch0 + j];
                dst[ch0 + j] = __float2half_rn(fmaxf(__fsub_rn(ahj, v[j]), 0.0f));
                dst[P.S_next + ch0 + j] = __float2half_rn(fmaxf(__fsub_rn(v[j], ahj), 0.0f));
              }
            }
          }
        };
#pragma unroll
        for (int c = 0; c < PF; c++) {
          const int j0 = 8 * (part + c * nparts);
          if (j0 < n_real) chunk(j0, ahp[c]);
        }
        for (int j0 = 8 * (part + PF * nparts); j0 < n_real; j0 += 8 * nparts) chunk(j0, nullptr);
        }
      } else if (EPI == 2) {
        // raw accumulator rows (the r_1-resolution half of the layer-0 gate convolution): fp32, 32-byte stores.
        // tcgen05.ld is warp-collective: every lane walks the loop, only the stores are predicated.
        {
          float *dst = P.g_out + (((long long)b * P.H + h) * P.W + w) * P.g_cols + nt * P.n_tile;
          const int n_real = P.g_cols - nt * P.n_tile < P.n_tile ? P.g_cols - nt * P.n_tile : P.n_tile;
          for (int j0 = 8 * part; j0 < n_real; j0 += 16 * nparts) {
            float va[8], vb[8];
            const int j1 = j0 + 8 * nparts;
            const bool has_b = j1 < n_real;   // warp-uniform
            tc_ld8(trow + j0, va);
            tc_ld8(trow + (has_b ? j1 : j0), vb);
            tc_ld_wait();
            if (valid) {
              asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + j0), "f"(va[0]),
                           "f"(va[1]), "f"(va[2]), "f"(va[3]), "f"(va[4]), "f"(va[5]), "f"(va[6]), "f"(va[7])
                           : "memory");
              if (has_b)
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + j1), "f"(vb[0]),
                             "f"(vb[1]), "f"(vb[2]), "f"(vb[3]), "f"(vb[4]), "f"(vb[5]), "f"(vb[6]), "f"(vb[7])
                             : "memory");
            }
          }
        }
      } else {
        // LSTM cell: columns [g*NCp + j], g = i,f,c,o.  c = f*C0 + i*tanh(.), r = o*tanh(c)
        const long long pix = (long long)h * P.W + w;
        const float *bm = P.bm + pix * 4 * P.R + nt * P.NC;
        const float *c0 = P.c0 + pix * P.R + nt * P.NC;
        // r for channels j0..j0+7 of this N tile -> fp16 into hv (or straight to the fp32 r_0 buffer at layer 0)
        auto lstm_chunk = [&](int j0, __half *hv) {
          float vi[8], vf[8], vc[8], vo[8];
          long long q0 = 0;
          if (edbg) q0 = clock64();
          tc_ld8(trow + 0 * P.NCp + j0, vi);
          tc_ld8(trow + 1 * P.NCp + j0, vf);
          tc_ld8(trow + 2 * P.NCp + j0, vc);
          tc_ld8(trow + 3 * P.NCp + j0, vo);
          tc_ld_wait();
          if (edbg) {
            const long long q1 = clock64();
            e_a += q1 - q0;
            q0 = q1;
          }
          if (!valid) return;
          float r[8];
          if (P.bm_packed) {
            const float *bp = P.bm + (pix * (P.R >> 3) + ((nt * P.NC + j0) >> 3)) * 32;
            float bq[32], cq[8];
            ldg256(bp, bq);
            ldg256(bp + 8, bq + 8);
            ldg256(bp + 16, bq + 16);
            ldg256(bp + 24, bq + 24);
            ldg256(c0 + j0, cq);
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float gi = hsig(__fadd_rn(vi[j], bq[j]));
              const float gf = hsig(__fadd_rn(vf[j], bq[8 + j]));
              const float gc = fast_tanh(__fadd_rn(vc[j], bq[16 + j]));
              const float go = hsig(__fadd_rn(vo[j], bq[24 + j]));
              const float c = __fadd_rn(__fmul_rn(gf, cq[j]), __fmul_rn(gi, gc));
              r[j] = __fmul_rn(go, fast_tanh(c));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              if (j0 + j < P.NC) {
                const float gi = hsig(__fadd_rn(vi[j], bm[0 * P.R + j0 + j]));
                const float gf = hsig(__fadd_rn(vf[j], bm[1 * P.R + j0 + j]));
                const float gc = fast_tanh(__fadd_rn(vc[j], bm[2 * P.R + j0 + j]));
                const float go = hsig(__fadd_rn(vo[j], bm[3 * P.R + j0 + j]));
                const float c = __fadd_rn(__fmul_rn(gf, c0[j0 + j]), __fmul_rn(gi, gc));
                r[j] = __fmul_rn(go, fast_tanh(c));
              } else {
                r[j] = 0.0f;
              }
            }
          }
          if (edbg) e_b += clock64() - q0;
          if (P.xr_out) {
#pragma unroll
            for (int j = 0; j < 8; j++) hv[j] = __float2half_rn(r[j]);
          } else {
            float *dst = P.r0_out + ((long long)b * P.H * P.W + pix) * P.R + nt * P.NC + j0;
            for (int j = 0; j < 8 && j0 + j < P.NC; j++) dst[j] = r[j];
          }
        };
        // two adjacent 8-channel chunks per pass, so that the up-sampled store is 32 bytes wide where alignment allows
        for (int jp = 16 * part; jp < P.NC; jp += 16 * nparts) {
          __align__(32) __half hv[16];
          const bool second = jp + 8 < P.NC;
          lstm_chunk(jp, hv);
          if (second) lstm_chunk(jp + 8, hv + 8);
          if (valid && P.xr_out && !P.xr_up) {
            // r at its own resolution (read by the r-resolution half of the layer-0 gate convolution)
            const int ch0 = nt * P.NC + jp;
            __half *dst = P.xr_out + ((long long)b * P.H * P.W + pix) * P.xr_cstride + P.xr_coff + ch0;
            if ((jp + 16 <= P.NC) && (((P.xr_coff + ch0) | P.xr_cstride) & 15) == 0) {
              const uint4 lo = *reinterpret_cast<const uint4 *>(hv), hi4 = *reinterpret_cast<const uint4 *>(hv + 8);
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(lo.x), "r"(lo.y),
                           "r"(lo.z), "r"(lo.w), "r"(hi4.x), "r"(hi4.y), "r"(hi4.z), "r"(hi4.w)
                           : "memory");
            } else {
              for (int j = 0; j < 16 && jp + j < P.NC; j++) dst[j] = hv[j];
            }
          } else if (valid && P.xr_out) {
            // nearest 2x up-sampling folded into the store: 4 destinations per source pixel
            const int H2 = P.H * 2, W2 = P.W * 2;
            const int ch0 = nt * P.NC + jp;
            const bool v32 = (jp + 16 <= P.NC) && (((P.xr_coff + ch0) | P.xr_cstride) & 15) == 0;
            const bool v16a = (jp + 8 <= P.NC) && (((P.xr_coff + ch0) | P.xr_cstride) & 7) == 0;
            const bool v16b = (jp + 16 <= P.NC) && (((P.xr_coff + ch0 + 8) | P.xr_cstride) & 7) == 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              __half *dst = P.xr_out + (((long long)b * H2 + 2 * h + (q >> 1)) * W2 + 2 * w + (q & 1)) * P.xr_cstride +
                            P.xr_coff + ch0;
              if (v32) {
                const uint4 lo = *reinterpret_cast<const uint4 *>(hv), hi4 = *reinterpret_cast<const uint4 *>(hv + 8);
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(lo.x), "r"(lo.y),
                             "r"(lo.z), "r"(lo.w), "r"(hi4.x), "r"(hi4.y), "r"(hi4.z), "r"(hi4.w)
                             : "memory");
              } else {
                if (v16a) *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hv);
                else for (int j = 0; j < 8 && jp + j < P.NC; j++) dst[j] = hv[j];
                if (second) {
                  if (v16b) *reinterpret_cast<uint4 *>(dst + 8) = *reinterpret_cast<const uint4 *>(hv + 8);
                  else for (int j = 8; j < 16 && jp + j < P.NC; j++) dst[j] = hv[j];
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO) mbar_arrive_cluster(tempty0 + 8 * a, 0);   // the leader waits for both CTAs' epilogues
        else mbar_arrive(tempty0 + 8 * a);
      }
      if (edbg) e_work += clock64() - ec0;
    }
    if (edbg) {   // TZ_CONV_DEBUG: epilogue warp 4 of this CTA: cycles waiting for accumulators / working, tiles done
      long long *o = P.dbg + 8 * blockIdx.x;
      o[4] = e_wait;
      o[5] = e_work;
      o[6] = e_tiles;
      o[7] = e_a;
      o[3] = e_b;
    }
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (TWO)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                   : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ small kernels
// prednet.py:274-277 at layer 0, t=0, written as fp16 into channels [0, 2C) of X_0.
__global__ void __launch_bounds__(256) e0_tc_kernel(const float *__restrict__ in, const float *__restrict__ p0,
                                                    __half *__restrict__ x0, long long total, long long frame, int C,
                                                    int cstride) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int c = (int)(id % C);
  long long pix = id / C;
  float a = in[id];
  float ah = p0[id % frame];
  x0[pix * cstride + c] = __float2half_rn(fmaxf(__fsub_rn(ah, a), 0.0f));
  x0[pix * cstride + C + c] = __float2half_rn(fmaxf(__fsub_rn(a, ah), 0.0f));
}
// The same for a 16-channel X_0 (the layer-0 split): one thread per pixel writes the whole 32-byte block -- the pad
// channels stay zero, and a full-sector store spares L2 the read-back that 2-byte stores into a sector cost
// (measured: 73 MB of DRAM reads per 100 frames for a kernel that reads 25 MB).
template <int C>
__global__ void __launch_bounds__(256) e0_px_kernel(const float *__restrict__ in, const float *__restrict__ p0,
                                                    __half *__restrict__ x0, long long npix, long long frame_px) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float *a = in + pix * C, *ah = p0 + (pix % frame_px) * C;
  __align__(16) __half ev[8];
#pragma unroll
  for (int j = 0; j < 8; j++) ev[j] = __float2half_rn(0.0f);
#pragma unroll
  for (int c = 0; c < C; c++) {
    const float av = a[c], hv = __ldg(ah + c);
    ev[c] = __float2half_rn(fmaxf(__fsub_rn(hv, av), 0.0f));
    ev[C + c] = __float2half_rn(fmaxf(__fsub_rn(av, hv), 0.0f));
  }
  const uint4 lo = *reinterpret_cast<const uint4 *>(ev);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %5, %5, %5};" ::"l"(x0 + pix * 16), "r"(lo.x), "r"(lo.y),
               "r"(lo.z), "r"(lo.w), "r"(0u)
               : "memory");
}

// BM [H*W, 4 gates, R] -> [H*W, R/8, 4 gates, 8]
__global__ void pack_bm_kernel(const float *__restrict__ src, float *__restrict__ dst, long long npix, int R) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * 4 * R) return;
  int j = (int)(i & 7);
  long long t = i >> 3;
  int g = (int)(t & 3);
  t >>= 2;
  int q = (int)(t % (R >> 3));
  long long pix = t / (R >> 3);
  dst[i] = src[(pix * 4 + g) * R + q * 8 + j];
}

// prednet.py:268-271 at layer 0: prediction = min(relu(conv3x3(r_0) + b), pixel_max), C -> C channels (C = 1 or 3).
// 16x16 output pixels per block; the 18x18xC input patch and the 9*C*C weights sit in shared memory.
// xe != nullptr: also stage the next step's layer-0 error units (what e0_tc_kernel would compute from `out`), so
// that a chained step (the usual case: frame k+1 is predicted from the prediction of frame k) skips that kernel.
template <int C>
struct Ahat0W {   // kernel parameter: the FMAs read the 9*C*C weights straight from the constant bank
  float w[9 * C * C];
  float b[C];
};

template <int C>
__global__ void __launch_bounds__(256) ahat0_kernel(const float *__restrict__ r0, const __grid_constant__ Ahat0W<C> wb,
                                                    float *__restrict__ out, int H, int W, float clip,
                                                    const float *__restrict__ p0, __half *__restrict__ xe, int cstride,
                                                    int epad) {
  constexpr int TW = 32, TH = 16, PT = TH / 8;   // output pixels per block; a warp is one image row, PT rows per thread
  __shared__ float tile[TH + 2][(TW + 2) * C];
  const int b = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const float *src = r0 + (long long)b * H * W * C;
  // rows of the patch are contiguous in memory ((TW+2)*C floats): consecutive threads read consecutive floats
  for (int yy = threadIdx.x / 32; yy < TH + 2; yy += 8) {
    const int gy = y0 + yy - 1;
    const bool rowok = gy >= 0 && gy < H;
    const float *row = src + ((long long)gy * W + (x0 - 1)) * C;
    for (int i = threadIdx.x & 31; i < (TW + 2) * C; i += 32) {
      const int gx = x0 - 1 + i / C;
      tile[yy][i] = (rowok && gx >= 0 && gx < W) ? row[i] : 0.0f;
    }
  }
  __syncthreads();
  const int tx = threadIdx.x & 31;
  const int x = x0 + tx;
#pragma unroll
  for (int pt = 0; pt < PT; pt++) {
    const int ty = (threadIdx.x >> 5) + 8 * pt;
    const int y = y0 + ty;
    if (x >= W || y >= H) continue;
    float acc[C];
#pragma unroll
    for (int co = 0; co < C; co++) acc[co] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
      for (int kx = 0; kx < 3; kx++)
#pragma unroll
        for (int ci = 0; ci < C; ci++) {
          const float v = tile[ty + ky][(tx + kx) * C + ci];
#pragma unroll
          for (int co = 0; co < C; co++) acc[co] = fmaf(v, wb.w[((ky * 3 + kx) * C + ci) * C + co], acc[co]);
        }
    const long long pix = ((long long)b * H + y) * W + x;
    float *dst = out + pix * C;
    __align__(16) __half ev[8];   // [relu(ahat - a) x C | relu(a - ahat) x C | zeros]: 2C <= 6 of the 8 lanes
#pragma unroll
    for (int j = 0; j < 8; j++) ev[j] = __float2half_rn(0.0f);
#pragma unroll
    for (int co = 0; co < C; co++) {
      const float a = fminf(fmaxf(acc[co] + wb.b[co], 0.0f), clip);
      dst[co] = a;
      if (xe) {   // prednet.py:274-277 of the NEXT step at layer 0, t = 0 (as e0_tc_kernel)
        const float ah = p0[((long long)y * W + x) * C + co];
        ev[co] = __float2half_rn(fmaxf(__fsub_rn(ah, a), 0.0f));
        ev[C + co] = __float2half_rn(fmaxf(__fsub_rn(a, ah), 0.0f));
      }
    }
    // One store per pixel covering the whole e block (channels 2C.. of it are padding that multiplies zero weights
    // and stays zero).  When the block is 16 channels = one 32-byte sector the store is 256 bits wide: a 16-byte
    // store leaves half a sector untouched and makes L2 read it back from DRAM first.
    if (xe) {
      const uint4 lo = *reinterpret_cast<const uint4 *>(ev);
      if (epad >= 16)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %5, %5, %5};" ::"l"(xe + pix * cstride), "r"(lo.x),
                     "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(0u)
                     : "memory");
      else
        *reinterpret_cast<uint4 *>(xe + pix * cstride) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------ layer-0 tail
// prednet.py:255-259 at layer 0 (gates + LSTM), :268-271 (A-hat_0 = prediction) and :274-277 of the NEXT step, in one
// kernel.  What is left of the layer-0 gate convolution once its up(r_1) half has moved to r_1's resolution (G,
// tc_create) is K = 9 taps x 2C error channels for 4C outputs.  As a tcgen05 implicit GEMM that is nine K16 slabs of
// which 10/16 are padding, one 44-cycle MMA each, behind an LSTM epilogue that waits for three global operands per
// pixel (measured 0.107 ms per 100 frames with the tensor pipe 8 % active); as plain FMAs it is 648 per pixel and
// issue-bound (0.122 ms).  Here a warp runs it as m16n8k16 register-level MMAs (fp16 x fp16 -> fp32, the numerics
// of the other convolutions): K = tap * 2C + channel is dense (54 of 64 for RGB, four K steps), the N = 16 columns
// are ordered so that lane t of a quad ends with the four gates of channel t for its two pixels, the bias map and
// G seed the accumulators, and the LSTM follows in registers.
// One block = 32 x 16 output pixels: e_0 on the 36 x 20 region (fp16 as stored, 16 bytes per pixel in shared
// memory), r_0 on the 34 x 18 region (the A-hat convolution needs a one-pixel halo of r_0, recomputed instead of
// exchanged; fp32), then the prediction (fp32 FMAs, weights in the constant bank), clipped, and the next step's
// error units.  X_0 is double-buffered: neighbouring blocks still read this step's error units while this block
// writes the next step's.
template <int C>
struct L0TailW {
  float wa[9 * C * C];           // A-hat kernel [tap][ci][co]
  float ba[C];
};
constexpr int L0_KSTEPS_MAX = 4;   // ceil(9 * 2C / 16) for C <= 3

__device__ __forceinline__ void mma_16816(float d[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// bfrag: [lane 32][K step][n tile 2][2] packed half2 B fragments of the e_0 slice of the gate kernel (tc_create);
// bm: the layer-0 bias map re-ordered to [pixel][channel][gate i, f, c, o] (one 16-byte load per pixel and channel)
#ifndef L0_MINB
#define L0_MINB 3
#endif
template <int C>
struct L0TailSmem {
  static constexpr int TW = 32, TH = 16, EW = TW + 4, EH = TH + 4, RW = TW + 2, RH = TH + 2;
  __half eh[EH][EW][8];          // e_0: 2C <= 8 channels per pixel (fp16 as stored)
  float4 a0[RW * RH * C];        // bias map + G per (pixel, channel): the accumulators' initial values [i, f, c, o]
  float c0[RW * RH * C];
  float rt[C][RH][RW];           // r_0
};

template <int C>
__global__ void __launch_bounds__(320, L0_MINB) l0_tail_kernel(const __half *__restrict__ xe_in, int cstride,
                                                         const float *__restrict__ bm, const float *__restrict__ c0,
                                                         const float *__restrict__ gr, const uint32_t *__restrict__ bfrag,
                                                         const __grid_constant__ L0TailW<C> wb, float *__restrict__ out,
                                                         int H, int W, float clip, const float *__restrict__ p0,
                                                         __half *__restrict__ xe_out) {
  using SM = L0TailSmem<C>;
  constexpr int TW = SM::TW, TH = SM::TH, EW = SM::EW, EH = SM::EH, RW = SM::RW, RH = SM::RH, NT = 320;
  constexpr int KS = (9 * 2 * C + 15) / 16;          // K steps
  constexpr int MT = (RW * RH + 15) / 16;            // 16-pixel M tiles of the r_0 region
  extern __shared__ __align__(16) uint8_t l0_smem[];
  SM &sm = *reinterpret_cast<SM *>(l0_smem);
  auto &rt = sm.rt;
  const int b = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- everything this block reads from global memory is requested up front (independent loads, many in flight
  // per thread); the MMA / LSTM loop below then runs out of shared memory.
  // (1) e_0 of the 36 x 20 region: one 16-byte load per pixel (the first 8 halves of its row; channels >= 2C are
  // zero in X_0 and multiply zero weights)
  for (int i = tid; i < EW * EH; i += NT) {
    const int ey = i / EW, ex = i - ey * EW;
    const int gy = y0 - 2 + ey, gx = x0 - 2 + ex;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      q = __ldg(reinterpret_cast<const uint4 *>(xe_in + (((long long)b * H + gy) * W + gx) * cstride));
    *reinterpret_cast<uint4 *>(&sm.eh[ey][ex][0]) = q;
  }
  // (2) per (pixel, channel) of the 34 x 18 region: bias map + G (this pixel's parity block of the r_1-resolution
  // convolution) and c(t=0).  (Measured: requesting all iterations' operands before the first store, or prefetching
  // them a pass ahead in registers, does not shorten the kernel -- it runs at ~70 % of the L1 data pipe's wavefront
  // rate, shared-memory operand reads of the MMAs and the sector-granular global accesses together.)
  for (int i = tid; i < RW * RH * C; i += NT) {
    const int p = i / C, ch = i - p * C;
    const int ry = p / RW, rx = p - ry * RW;
    const int gy = y0 - 1 + ry, gx = x0 - 1 + rx;
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float cv = 0.0f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const long long pix = (long long)gy * W + gx;
      const float4 gq = __ldg(reinterpret_cast<const float4 *>(
                                  gr + ((((long long)b * (H >> 1) + (gy >> 1)) * (W >> 1) + (gx >> 1)) * 4 +
                                        ((gy & 1) * 2 + (gx & 1))) * (4 * C)) + ch);
      const float4 bq = __ldg(reinterpret_cast<const float4 *>(bm) + pix * C + ch);   // [pixel][channel][i, f, c, o]
      cv = __ldg(c0 + pix * C + ch);
      v = make_float4(__fadd_rn(bq.x, gq.x), __fadd_rn(bq.y, gq.y), __fadd_rn(bq.z, gq.z), __fadd_rn(bq.w, gq.w));
    }
    sm.a0[i] = v;
    sm.c0[i] = cv;
  }
  // this lane's B fragments and the shared-memory offsets of its A columns (k = 16 s + 2 t + 8 j, k = tap * 2C + ci)
  const int g = lane >> 2, t = lane & 3;
  uint32_t bf[KS][2][2];
  int aoff[KS][2];
#pragma unroll
  for (int s_ = 0; s_ < KS; s_++) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const uint2 v = __ldg(reinterpret_cast<const uint2 *>(bfrag + ((lane * L0_KSTEPS_MAX + s_) * 2 + q) * 2));
      bf[s_][q][0] = v.x;
      bf[s_][q][1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int k = 16 * s_ + 2 * t + 8 * j;
      const int tap = k / (2 * C), ci = k - tap * 2 * C;
      aoff[s_][j] = (k < 9 * 2 * C) ? (((tap / 3) * EW + (tap % 3)) * 16 + ci * 2) : -1;
    }
  }
  __syncthreads();
  // ---- r_0 on the 34 x 18 region, 16 pixels per warp pass
  const char *ebase = reinterpret_cast<const char *>(&sm.eh[0][0][0]);
  for (int mt = warp; mt < MT; mt += NT / 32) {   // warp-uniform: mma.sync is warp-collective
    int pr[2], pp[2];
    bool ok[2];
    float acc[2][4], cprev[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      int p = mt * 16 + g + 8 * r;
      const bool inreg = p < RW * RH;
      if (!inreg) p = RW * RH - 1;
      pp[r] = p;
      const int ry = p / RW, rx = p - ry * RW;
      pr[r] = (ry * EW + rx) * 16;
      const int py = y0 - 1 + ry, px = x0 - 1 + rx;
      ok[r] = inreg && py >= 0 && py < H && px >= 0 && px < W;
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      cprev[r] = 0.0f;
      if (t < C) {
        v = sm.a0[p * C + t];
        cprev[r] = sm.c0[p * C + t];
      }
      acc[0][2 * r + 0] = v.x;   // n tile 0: columns (2t, 2t+1) = gates i, f of channel t
      acc[0][2 * r + 1] = v.y;
      acc[1][2 * r + 0] = v.z;   // n tile 1: gates c, o
      acc[1][2 * r + 1] = v.w;
    }
#pragma unroll
    for (int s_ = 0; s_ < KS; s_++) {
      uint32_t a[4];
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int r = 0; r < 2; r++)
          a[2 * j + r] = aoff[s_][j] >= 0 ? *reinterpret_cast<const uint32_t *>(ebase + pr[r] + aoff[s_][j]) : 0u;
      mma_16816(acc[0], a[0], a[1], a[2], a[3], bf[s_][0][0], bf[s_][0][1]);
      mma_16816(acc[1], a[0], a[1], a[2], a[3], bf[s_][1][0], bf[s_][1][1]);
    }
    if (t < C) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (mt * 16 + g + 8 * r < RW * RH) {
          float rv = 0.0f;   // outside the image: Keras' zero padding of the A-hat convolution
          if (ok[r]) {
            const float gi = hsig(acc[0][2 * r]), gf = hsig(acc[0][2 * r + 1]);
            const float gc = fast_tanh(acc[1][2 * r]), go = hsig(acc[1][2 * r + 1]);
            const float c = __fadd_rn(__fmul_rn(gf, cprev[r]), __fmul_rn(gi, gc));
            rv = __fmul_rn(go, fast_tanh(c));
          }
          const int ry = pp[r] / RW, rx = pp[r] - ry * RW;
          rt[t][ry][rx] = rv;
        }
      }
    }
  }
  __syncthreads();
  // ---- prediction = min(relu(conv3x3(r_0) + b), pixel_max) and the next step's error units
  for (int i = tid; i < TW * TH; i += NT) {
    const int ty = i / TW, tx = i - ty * TW;
    const int y = y0 + ty, x = x0 + tx;
    if (x >= W || y >= H) continue;
    float acc[C];
#pragma unroll
    for (int co = 0; co < C; co++) acc[co] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
      for (int kx = 0; kx < 3; kx++)
#pragma unroll
        for (int ci = 0; ci < C; ci++) {
          const float v = rt[ci][ty + ky][tx + kx];
#pragma unroll
          for (int co = 0; co < C; co++) acc[co] = fmaf(v, wb.wa[((ky * 3 + kx) * C + ci) * C + co], acc[co]);
        }
    const long long pix = ((long long)b * H + y) * W + x;
    float *dst = out + pix * C;
    __align__(16) __half ev[8];
#pragma unroll
    for (int j = 0; j < 8; j++) ev[j] = __float2half_rn(0.0f);
#pragma unroll
    for (int co = 0; co < C; co++) {
      const float a = fminf(fmaxf(acc[co] + wb.ba[co], 0.0f), clip);
      dst[co] = a;
      const float ah = __ldg(p0 + ((long long)y * W + x) * C + co);
      ev[co] = __float2half_rn(fmaxf(__fsub_rn(ah, a), 0.0f));
      ev[C + co] = __float2half_rn(fmaxf(__fsub_rn(a, ah), 0.0f));
    }
    const uint4 lo = *reinterpret_cast<const uint4 *>(ev);
    if (cstride >= 16)   // the whole 32-byte sector of the e_0 block (its upper half is padding and stays zero)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %5, %5, %5};" ::"l"(xe_out + pix * cstride), "r"(lo.x),
                   "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(0u)
                   : "memory");
    else
      *reinterpret_cast<uint4 *>(xe_out + pix * cstride) = lo;
  }
}

__global__ void f32_to_f16_kernel(const float *__restrict__ src, __half *__restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

// ------------------------------------------------------------------------------------------------ host side
struct ConvTc {
  CUtensorMap tmA, tmB;
  ConvArgs args;
  int epi;
  uint32_t smem_bytes;
  __half *wpack;
};

}  // namespace tz

struct TcState {
  int L;
  __half *X[TZ_MAX_LAYERS];   // X_l: [maxB, H_l, W_l, cx[l]] fp16: [e_l | up(r_{l+1}) | zero pad]
  int cx[TZ_MAX_LAYERS];
  int epad[TZ_MAX_LAYERS];    // channel offset of the up(r_{l+1}) block: 2*S_l rounded up to 16 (32-byte stores)
  float *r0;                  // [maxB, H_0, W_0, R_0] fp32
  // layer 0 with the up(r_1) half of its gate convolution evaluated at r_1's resolution (use_gr):
  bool use_gr;
  __half *rlow;               // r_1 at its own resolution: [maxB, H_1, W_1, cxr] fp16 (channels >= R_1 stay zero)
  int cxr;
  float *gr;                  // G: [maxB, H_1, W_1, 4 parities, R_0, 4 gates] fp32
  tz::ConvTc rconv0;          // r_1 -> G (raw epilogue)
  __half *X0b;                // second X_0 buffer: the layer-0 tail kernel reads one and writes the other
  tz::ConvTc aconv0_alt;      // a_0 reading X0b
  int x0_cur;                 // which of X[0] / X0b holds the current step's error units
  uint32_t *l0_bfrag;         // layer-0 tail kernel: B fragments of the e_0 slice of the gate kernel (per lane)
  float *l0_bm;               // BM_0 as [pixel][channel][gate]
  tz::ConvTc aconv[TZ_MAX_LAYERS];   // l = 0..L-2
  tz::ConvTc gconv[TZ_MAX_LAYERS];   // l = 0..L-1
  int sm_count;
  float ahat0_w[9 * 3 * 3], ahat0_b[3];   // host copy of the layer-0 A-hat kernel (C = R_0 = S_0 <= 3), passed by value
};

namespace tz {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }
static int pick_kc(int c) { return (c % 64 == 0) ? 64 : (c % 32 == 0) ? 32 : 16; }
static int largest_divisor_le(int n, int cap) {
  for (int d = cap < n ? cap : n; d >= 1; d--)
    if (n % d == 0) return d;
  return 1;
}

// tile = 2^tb x 2^th x 2^tw pixels (128 in total), chosen to minimise padded work at this resolution.
static void pick_tile(int H, int W, bool pool, int *tw_log, int *th_log, int *tb_log) {
  double best = 1e30;
  for (int a = pool ? 1 : 0; a <= (pool ? 4 : 7); a++)      // TW <= 16 when the epilogue pools inside a warp
    for (int b = pool ? 1 : 0; a + b <= 7; b++) {
      int TW = 1 << a, TH = 1 << b, TB = 128 / (TW * TH);
      if (TW > 256 || TH > 256 || TB > 128) continue;
      if (pool && TW * 2 > 32) continue;
      double cost = (double)((W + TW - 1) / TW) * ((H + TH - 1) / TH) / TB + 1e-3 * (7 - a) + 1e-4 * TB;
      if (cost < best) {
        best = cost;
        *tw_log = a;
        *th_log = b;
        *tb_log = 7 - a - b;
      }
    }
}

// cmap[i] = input-channel row of the fp32 kernel that multiplies channel i of the activation buffer X (-1: none,
// the packed weight is zero).  The conv reads channels [0, round_up(cmap.size(), 16)) of X.
static int make_conv(tz_prednet *h, ConvTc *c, int epi, int l, __half *X, int cx, const std::vector<int> &cmap,
                     const std::vector<float> &wsrc /*fp32 [3][3][cin_w][cout_w]*/, int cin_w, int cout_w,
                     int n_real /*output channels or R*/) {
  const int cin_real = (int)cmap.size();
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TZ_ECUDA;
  }
  memset(c, 0, sizeof(*c));
  ConvArgs &A = c->args;
  c->epi = epi;
  A.H = h->H[l];
  A.W = h->W[l];
  pick_tile(A.H, A.W, epi == 0, &A.tw_log, &A.th_log, &A.tb_log);
  A.tiles_w = (A.W + (1 << A.tw_log) - 1) >> A.tw_log;
  A.tiles_h = (A.H + (1 << A.th_log) - 1) >> A.th_log;
  A.cin_pad = round_up(cin_real, 16);
  if (A.cin_pad > cx) {
    set_error("internal: conv reads %d channels of a %d-channel buffer", A.cin_pad, cx);
    return TZ_EINVAL;
  }
  A.KC = pick_kc(A.cin_pad);
  A.kchunks = A.cin_pad / A.KC;
  A.ksteps = A.KC / 16;
  int rows_total;
  int n_unit = 0;   // output channels (A path) or R channels (R path) per N tile
  if (epi == 1) {
    A.R = n_real;
    n_unit = largest_divisor_le(n_real, 64);
    // A/B switch for the im2col-fed gate convolution (image width not a multiple of 8: gates3 of the 128x160 net):
    // R channels per N tile.  64 (N = 256) leaves 750 tile units for 148 CTAs; 48 (N = 192) gives 1000 smaller ones.
    const char *gu = getenv("TZ_GATE_UNIT");
    if (gu && (A.W % 8) != 0) {
      const int u = atoi(gu);
      if (u > 0 && u <= 64 && n_real % u == 0) n_unit = u;
    }
  } else {
    n_unit = largest_divisor_le(n_real, 256);
  }
  // A/B switch, read at create time: TZ_HALO=0 im2col feeding only, TZ_HALO=1 no CTA pairs.
  const char *halo_env = getenv("TZ_HALO");
  // ---- A-path convs whose stationary weights would need more than the 112 KB that leave room for four halo stages
  // (a1 of the (3,48,96,192) net: 166 KB) run faster as a CTA pair with streamed weights, even with the input
  // channels rounded up to 64 (measured: 0.080 vs 0.086 ms); smaller ones do not (a0 0.158 vs 0.101 ms).
  {
    const int c64 = round_up(cin_real, 64);
    const int ntile_c = round_up(n_unit, 16);
    const bool big_a = epi == 0 && !halo_env && (size_t)ntile_c * 9u * (size_t)A.cin_pad * 2u > 112u * 1024u;
    if (big_a && c64 <= cx && (A.W % 8) == 0 && (A.H % 16) == 0 && ntile_c <= 256) {
      A.halo = 4;
      A.cin_pad = c64;
      A.KC = 64;
      A.kchunks = c64 / 64;
      A.ksteps = 4;
      A.tw_log = 3;
      A.th_log = 4;
      A.tb_log = 0;
      A.tiles_w = (A.W + 7) >> 3;
      A.tiles_h = (A.H + 15) >> 4;
    }
  }
  // ---- halo + stationary-weights mode: small weight matrices only (they must fit in shared memory next to the
  // halo tiles), image width a multiple of the 8-pixel tile width.
  if (!A.halo) {
    const bool want = !(halo_env && halo_env[0] == '0');
    const int cin64 = A.cin_pad;            // channels read: cin rounded up to 16, walked in chunks of KC
    // Stationary weights may take what two halo stages leave free: one wide N tile halves the MMA count of a split
    // one (measured, a1 of the (3,48,96,192) net: N = 96 in one tile 0.098 ms, two tiles of 48 0.129 ms).
    const uint32_t halo_stage = (16u * 18u * 2u * (uint32_t)A.KC + 1023u) & ~1023u;
    const uint32_t budget = 224u * 1024u - 2u * halo_stage;
    if (want && cin64 <= cx && (A.W % 8) == 0) {
      for (int split = 1; split <= 2; split++) {
        if (n_real % split) continue;
        const int unit = n_real / split;
        if (epi == 1 && unit > 64) continue;
        if (epi != 1 && split > 1 && (unit % 16) != 0) continue;
        const int ntile = (epi == 1) ? round_up(4 * round_up(unit, 8), 16) : round_up(unit, 16);
        const uint32_t wbytes = 9u * (uint32_t)(cin64 / A.KC) * (((uint32_t)ntile * 2u * (uint32_t)A.KC + 1023u) & ~1023u);
        if (wbytes > (split == 1 ? budget : 112u * 1024u) || ntile > 256) continue;
        A.halo = 1;
        n_unit = unit;
        A.tw_log = 3;
        A.th_log = 4;
        A.tb_log = 0;
        A.tiles_w = (A.W + 7) >> 3;
        A.tiles_h = (A.H + 15) >> 4;
        break;
      }
    }
  }
  // ---- CTA-pair mode (cta_group::2): large weight matrices.  Clusters of two CTAs run M256 MMAs: each CTA holds the
  // halo box of its own 8x16 M tile and HALF of every streamed weight block, so per SM both the TMA fill and the
  // tensor core's shared-memory reads of B are halved (shared-memory bandwidth is what bounds the one-CTA modes).
  if (!A.halo) {
    const bool want2 = !halo_env;
    const int ntile_c = (epi == 1) ? round_up(4 * round_up(n_unit, 8), 16) : round_up(n_unit, 16);
    if (want2 && epi != 2 && (A.cin_pad % 64) == 0 && A.cin_pad <= cx && (A.W % 8) == 0 && (A.H % 16) == 0 && ntile_c <= 256 &&
        (ntile_c % 16) == 0) {
      A.halo = 4;
      A.KC = 64;
      A.kchunks = A.cin_pad / 64;
      A.ksteps = 4;
      A.tw_log = 3;
      A.th_log = 4;
      A.tb_log = 0;
      A.tiles_w = (A.W + 7) >> 3;
      A.tiles_h = (A.H + 15) >> 4;
    }
  }
  A.tile_h = 1 << A.th_log;
  A.ksteps_last = A.ksteps;
  if (A.halo == 4) {
    const int last = round_up(cin_real, 16) - (A.kchunks - 1) * 64;   // channels of the last block that carry weights
    if (last > 0 && last <= 32) A.ksteps_last = 2;
  }
  if (epi == 1) {
    A.NC = n_unit;
    A.NCp = round_up(A.NC, 8);
    A.n_tile = round_up(4 * A.NCp, 16);
    A.n_tiles_n = n_real / A.NC;
  } else {
    A.n_tile = round_up(n_unit, 16);
    A.n_tiles_n = n_real / n_unit;
    if (A.n_tiles_n > 1 && (n_unit % 16) != 0) {
      set_error("unsupported channel count %d for the tensor-core path", n_real);
      return TZ_EINVAL;
    }
    // pooling epilogue with 8-byte stores (conv_tc_kernel, EPI 0): whole groups of 16 channels per tile, even image
    // and 8 x 16 or larger power-of-two tiles, a 32-byte aligned e block in X_{l+1}
    A.perm16 = epi == 0 && (n_unit % 16) == 0 && (A.H % 2) == 0 && (A.W % 2) == 0 && A.tw_log >= 1 && A.th_log >= 1 &&
               !getenv("TZ_NO_PERM16");
  }
  rows_total = A.n_tiles_n * A.n_tile;
  const int Ktot = 9 * A.cin_pad;
  // ---- pack weights: row n (tile-major; gates interleaved per tile), K-major, k = tap*cin_pad + c
  std::vector<float> wp((size_t)rows_total * Ktot, 0.0f);
  for (int nt = 0; nt < A.n_tiles_n; nt++)
    for (int r = 0; r < A.n_tile; r++) {
      int src_col = -1;
      if (epi == 1) {
        int g = r / A.NCp, j = r - g * A.NCp;
        if (g < 4 && j < A.NC) src_col = g * n_real + nt * A.NC + j;
      } else {
        int per = n_real / A.n_tiles_n;
        if (r < per) {
          int ch = r;
          if (A.perm16) {   // column r of the tile holds channel: chunk A (r % 16 < 8) -> 2 of every 4, chunk B the others
            const int k16 = r >> 4, c16 = r & 15, c = c16 & 7;
            ch = 16 * k16 + 2 * (c & ~1) + (c & 1) + (c16 >= 8 ? 2 : 0);
          }
          src_col = nt * per + ch;
        }
      }
      if (src_col < 0) continue;
      float *dst = wp.data() + (size_t)(nt * A.n_tile + r) * Ktot;
      for (int tap = 0; tap < 9; tap++)
        for (int ci = 0; ci < cin_real; ci++)
          if (cmap[ci] >= 0) dst[tap * A.cin_pad + ci] = wsrc[((size_t)tap * cin_w + cmap[ci]) * cout_w + src_col];
    }
  float *tmp = nullptr;
  TZ_CHECK_CUDA(cudaMalloc(&tmp, wp.size() * sizeof(float)));
  c->wpack = (__half *)dev_alloc(h, wp.size() * sizeof(__half));
  if (!c->wpack) {
    cudaFree(tmp);
    return TZ_ENOMEM;
  }
  cudaMemcpy(tmp, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice);
  f32_to_f16_kernel<<<(unsigned)((wp.size() + 255) / 256), 256>>>(tmp, c->wpack, (long long)wp.size());
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(tmp);
  if (e != cudaSuccess) {
    set_error("weight packing failed: %s", cudaGetErrorString(e));
    return TZ_ECUDA;
  }
  // ---- shared-memory plan
  const uint32_t row_bytes = (uint32_t)A.KC * 2;
  const uint32_t a_bytes = 128u * row_bytes, b_bytes = (uint32_t)A.n_tile * row_bytes;
  A.a_stride = (a_bytes + 1023u) & ~1023u;
  A.stage_stride = A.a_stride + ((b_bytes + 1023u) & ~1023u);
  A.tx_bytes = a_bytes + b_bytes;
  int stages = (int)((200u * 1024u) / A.stage_stride);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) {
    set_error("internal: pipeline stage of %u bytes does not fit", A.stage_stride);
    return TZ_EINVAL;
  }
  A.stages = stages;
  c->smem_bytes = (uint32_t)stages * A.stage_stride + 1024u;
  if (A.halo == 4) {
    A.a_tx = 10u * 18u * row_bytes;     // 18 image rows x 10 pixels x KC channels fp16 (this CTA's M tile + halo)
    A.a_slot = (A.a_tx + 1023u) & ~1023u;
    A.a_stride = 0;
    A.b_block = (((uint32_t)(A.n_tile / 2) * row_bytes) + 1023u) & ~1023u;   // this CTA's half of one tap's weight block
    A.tx_bytes = 3u * (uint32_t)(A.n_tile / 2) * row_bytes;                  // a stage holds one filter row (3 taps)
    A.stage_stride = 3u * A.b_block;
    stages = (int)((226u * 1024u - 1024u - 2u * A.a_slot) / A.stage_stride);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    A.stages = stages;
    c->smem_bytes = 2u * A.a_slot + (uint32_t)stages * A.stage_stride + 1024u;
  } else if (A.halo) {
    A.b_block = ((uint32_t)A.n_tile * row_bytes + 1023u) & ~1023u;
    A.b_region = 9u * (uint32_t)A.kchunks * A.b_block;
    A.a_stride = 0;
    A.stage_stride = (16u * 18u * row_bytes + 1023u) & ~1023u;   // one halo tile: 18 image rows x 16 pixels x KC channels
    A.tx_bytes = 16u * 18u * row_bytes;
    stages = (int)((224u * 1024u - A.b_region) / A.stage_stride);
    if (stages > 4) stages = 4;
    if (stages < 2) {
      set_error("internal: halo mode does not fit (weights %u bytes)", A.b_region);
      return TZ_EINVAL;
    }
    A.stages = stages;
    c->smem_bytes = A.b_region + (uint32_t)stages * A.stage_stride + 1024u;
  }
  {
    // work items per tile that the epilogue warps share out: 8-channel chunks (A path) / 16-channel pairs (R path)
    const int nchunks = (epi == 1) ? (A.NC + 15) / 16 : (A.n_tile + 7) / 8;
    A.epi_warps = nchunks >= 3 ? 12 : nchunks >= 2 ? 8 : 4;
    A.epi_groups = 1;
    A.acc_stages = 2;
    A.acc_stride = 256;
    if (A.halo != 4 && A.n_tile <= 64) {   // narrow tiles: the epilogue is latency-bound, keep three of them in flight
      A.epi_warps = 12;
      A.epi_groups = 3;
      A.acc_stages = 6;
      A.acc_stride = 64;
    }
  }
  {
    auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)((0x100000000ULL + (uint64_t)d - 1) / (uint64_t)d); };
    A.m_tiles_w = magic(A.tiles_w);
    A.m_tiles_h = magic(A.tiles_h);
    A.m_tiles_n = magic(A.n_tiles_n);
  }
  {
    // Two issuers take alternate tiles.  mbarrier waits are by phase PARITY, so a thread may only wait on a barrier
    // whose previous phase it has itself seen complete: every halo slot and every accumulator must always belong to
    // the same issuer, i.e. one chunk per tile and even ring sizes (otherwise a wait one full phase ahead returns
    // at once -- seen as a launch failure on a1, three chunks per tile, when this was unconditional).
    const char *env = getenv("TZ_MMA_ISSUERS");   // A/B switch
    A.mma_issuers = (A.halo == 1 && A.kchunks == 1 && (A.stages % 2) == 0 && (A.acc_stages % 2) == 0 &&
                     !(env && env[0] == '1')) ? 2 : 1;
  }
  // ---- descriptors
  const CUtensorMapSwizzle swz = A.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : A.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const uint32_t layout_type = A.KC == 64 ? 2u : A.KC == 32 ? 4u : 6u;   // UMMA SWIZZLE_128B / 64B / 32B
  const uint32_t sbo = 8u * row_bytes;                                   // 8-row core-matrix group stride
  A.desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
  A.idesc = (1u << 4) | ((uint32_t)(A.n_tile >> 3) << 17) | (((A.halo == 4 ? 256u : 128u) >> 4) << 24);   // f32 acc, f16 x f16, K-major
  {
    cuuint64_t dims[4] = {(cuuint64_t)cx, (cuuint64_t)A.W, (cuuint64_t)A.H, (cuuint64_t)h->cfg.max_batch};
    cuuint64_t strides[3] = {(cuuint64_t)cx * 2, (cuuint64_t)cx * 2 * A.W, (cuuint64_t)cx * 2 * A.W * A.H};
    cuuint32_t box[4] = {(cuuint32_t)A.KC, 1u << A.tw_log, 1u << A.th_log, 1u << A.tb_log};
    if (A.halo) {   // (tile height + 2) image rows of 16 pixels each (8-wide tile + halo, padded to a 16-pixel pitch)
      box[1] = (A.halo == 4) ? 10 : 16;
      box[2] = (cuuint32_t)A.tile_h + 2;
      box[3] = 1;
    }
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&c->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, X, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A, layer %d) failed: %d", l, (int)r);
      return TZ_ECUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows_total};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)A.KC, (cuuint32_t)(A.halo == 4 ? A.n_tile / 2 : A.n_tile)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&c->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c->wpack, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B, layer %d) failed: %d", l, (int)r);
      return TZ_ECUDA;
    }
  }
  return TZ_OK;
}

int tc_create(tz_prednet *h, const std::vector<std::vector<float>> &wg_host) {
  const int L = h->L;
  TcState *T = new TcState();
  memset(T, 0, sizeof(*T));
  h->tc = T;
  T->L = L;
  T->sm_count = sm_count();
  const int mb = h->cfg.max_batch;
  // Layer 0 is the widest image and the narrowest convolution (N = 4 R_0 = 12 columns for RGB): an M128 x K16 MMA
  // costs 44 cycles however narrow N is (it reads 4 KB of A from shared memory), so its gate convolution is bound by
  // the NUMBER of K16 slabs, and 27 of the 36 slabs per tile come from up(r_1).  A 3x3 convolution over a
  // nearest-neighbour up-sampled image is, for each of the four output parities, a convolution over the
  // low-resolution image with pre-summed taps (prednet.py:250-258 with UpSampling2D): those 27 slabs per 128 pixels
  // become one tcgen05 convolution at r_1's resolution with 4 x 4R_0 output columns (36 slabs per 512 pixels,
  // `rconv0`, raw fp32 accumulators G).  What remains of layer 0 -- the e_0 half of the gates (K = 54), + bias map
  // + G, the LSTM, the A-hat_0 convolution and the next step's error units -- is one kernel (l0_tail_kernel).  r_1
  // is never written up-sampled and X_0 holds only the 32-byte e_0 block (double-buffered).  Used when S_0 = R_0 is 1
  // or 3 (every net the reference builds); other nets keep the generic kernels.
  T->use_gr = L >= 2 && h->R[0] == h->S[0] && (h->S[0] == 3 || h->S[0] == 1) && !getenv("TZ_NO_GR");
  for (int l = 0; l < L; l++) {
    TZ_REQUIRE(h->H[l] % 2 == 0 || l == L - 1, "tensor-core path: odd layer height");
    T->epad[l] = round_up(2 * h->S[l], 16);   // 32-byte aligned r_up block
    T->cx[l] = round_up(T->epad[l] + ((l < L - 1 && !(l == 0 && T->use_gr)) ? h->R[l + 1] : 0), 16);
    size_t bytes = (size_t)mb * h->H[l] * h->W[l] * T->cx[l] * sizeof(__half);
    T->X[l] = (__half *)dev_alloc(h, bytes);
    if (!T->X[l]) return TZ_ENOMEM;
    TZ_CHECK_CUDA(cudaMemset(T->X[l], 0, bytes));   // pad channels multiply zero weights: they must stay finite
  }
  if (!T->use_gr) {   // (with the layer-0 split r_0 never leaves the chip)
    T->r0 = (float *)dev_alloc(h, (size_t)mb * h->H[0] * h->W[0] * h->R[0] * sizeof(float));
    if (!T->r0) return TZ_ENOMEM;
  }
  if (T->use_gr) {
    T->cxr = round_up(h->R[1], 64);
    const size_t px1 = (size_t)mb * h->H[1] * h->W[1];
    T->rlow = (__half *)dev_alloc(h, px1 * T->cxr * sizeof(__half));
    T->gr = (float *)dev_alloc(h, px1 * 16 * h->R[0] * sizeof(float));
    if (!T->rlow || !T->gr) return TZ_ENOMEM;
    TZ_CHECK_CUDA(cudaMemset(T->rlow, 0, px1 * T->cxr * sizeof(__half)));
    const size_t xb = (size_t)mb * h->H[0] * h->W[0] * T->cx[0] * sizeof(__half);
    T->X0b = (__half *)dev_alloc(h, xb);
    if (!T->X0b) return TZ_ENOMEM;
    TZ_CHECK_CUDA(cudaMemset(T->X0b, 0, xb));
  }
  if (h->R[0] == h->S[0] && (h->S[0] == 3 || h->S[0] == 1)) {
    const int C = h->S[0];
    TZ_CHECK_CUDA(cudaMemcpy(T->ahat0_w, h->w_ahat[0], sizeof(float) * 9 * C * C, cudaMemcpyDeviceToHost));
    TZ_CHECK_CUDA(cudaMemcpy(T->ahat0_b, h->b_ahat[0], sizeof(float) * C, cudaMemcpyDeviceToHost));
  }
  TZ_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  TZ_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  TZ_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  TZ_CHECK_CUDA(cudaFuncSetAttribute(l0_tail_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(L0TailSmem<3>)));
  TZ_CHECK_CUDA(cudaFuncSetAttribute(l0_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(L0TailSmem<1>)));
  TZ_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  TZ_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));

  for (int l = 0; l < L; l++) {
    // gate conv: reads all of X_l = [e_l | up(r_{l+1})]; the r_{t-1} slice of the kernel is hoisted into BM_l
    const bool r_in_x = l < L - 1 && !(l == 0 && T->use_gr);   // X_l carries the up-sampled r_{l+1} block
    std::vector<int> gmap(T->epad[l] + (r_in_x ? h->R[l + 1] : 0), -1);
    for (int i = 0; i < 2 * h->S[l]; i++) gmap[i] = h->R[l] + i;                                  // e_l
    if (r_in_x)
      for (int i = 0; i < h->R[l + 1]; i++) gmap[T->epad[l] + i] = h->R[l] + 2 * h->S[l] + i;    // up(r_{l+1})
    int rc = TZ_OK;
    const bool tail0 = l == 0 && T->use_gr;   // layer-0 gates: r_1 half on the tensor core (rconv0), the rest in l0_tail_kernel
    if (!tail0) {
      rc = make_conv(h, &T->gconv[l], 1, l, T->X[l], T->cx[l], gmap, wg_host[l], h->cin_g[l], 4 * h->R[l], h->R[l]);
      if (rc) return rc;
    }
    ConvArgs &G = T->gconv[l].args;
    G.bm = h->BM[l];
    G.bm_packed = 0;
    G.c0 = h->C0[l];
    if (!tail0 && G.NC % 8 == 0 && h->R[l] % 8 == 0) {
      const long long npix = (long long)h->H[l] * h->W[l];
      float *bp = (float *)dev_alloc(h, (size_t)npix * 4 * h->R[l] * sizeof(float));
      if (!bp) return TZ_ENOMEM;
      const long long n = npix * 4 * h->R[l];
      pack_bm_kernel<<<(unsigned)((n + 255) / 256), 256>>>(h->BM[l], bp, npix, h->R[l]);
      TZ_CHECK_CUDA(cudaDeviceSynchronize());
      G.bm = bp;
      G.bm_packed = 1;
    }
    G.xr_up = 1;
    if (l == 1 && T->use_gr) {
      G.xr_out = T->rlow;
      G.xr_cstride = T->cxr;
      G.xr_coff = 0;
      G.xr_up = 0;
    } else if (l > 0) {
      G.xr_out = T->X[l - 1];
      G.xr_cstride = T->cx[l - 1];
      G.xr_coff = T->epad[l - 1];
    } else {
      G.xr_out = nullptr;
      G.r0_out = T->r0;
    }
    if (l == 0 && T->use_gr) {
      // the up(r_1) half of the layer-0 gates at r_1's resolution.  Output parity (py, px), low-resolution offset
      // (oy, ox): W_eff = sum of the taps (dy, dx) with floor((py + dy) / 2) == oy and floor((px + dx) / 2) == ox
      // (py = 0: dy = -1 -> oy = -1, dy = 0, 1 -> oy = 0;  py = 1: dy = -1, 0 -> oy = 0, dy = 1 -> oy = 1).  Keras'
      // zero padding acts on the up-sampled image; its rows -1 and 2H_1 are the low-resolution rows -1 and H_1, so
      // TMA's out-of-bounds zero fill reproduces it exactly.  Columns: [parity][channel][gate i,f,c,o].
      const int R0 = h->R[0], R1 = h->R[1], cols = 16 * R0, cin = h->cin_g[0], coff = R0 + 2 * h->S[0];
      std::vector<float> weff((size_t)9 * R1 * cols, 0.0f);
      for (int py = 0; py < 2; py++)
        for (int px = 0; px < 2; px++)
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              const int oy = (py + dy + 2) / 2 - 1, ox = (px + dx + 2) / 2 - 1;   // floor((p + d) / 2)
              const int tsrc = (dy + 1) * 3 + (dx + 1), tdst = (oy + 1) * 3 + (ox + 1);
              for (int ci = 0; ci < R1; ci++)
                for (int g = 0; g < 4; g++)
                  for (int ch = 0; ch < R0; ch++)
                    weff[((size_t)tdst * R1 + ci) * cols + ((py * 2 + px) * R0 + ch) * 4 + g] +=
                        wg_host[0][((size_t)tsrc * cin + coff + ci) * 4 * R0 + g * R0 + ch];
            }
      std::vector<int> rmap(T->cxr, -1);
      for (int i = 0; i < R1; i++) rmap[i] = i;
      rc = make_conv(h, &T->rconv0, 2, 1, T->rlow, T->cxr, rmap, weff, R1, cols, cols);
      if (rc) return rc;
      T->rconv0.args.g_out = T->gr;
      T->rconv0.args.g_cols = cols;
      // e_0 slice of the gate kernel as m16n8k16 B fragments (l0_tail_kernel): k = tap * 2 S_0 + ci; column
      // 8 q + 2 t' + u = gate (2 q + u) of channel t'.  Lane (g, t) holds b0 = (k, k + 1) and b1 = (k + 8, k + 9)
      // with k = 16 s + 2 t, for column n = 8 q + g.
      {
        const int KC2 = 2 * h->S[0];
        std::vector<uint32_t> bf((size_t)32 * L0_KSTEPS_MAX * 2 * 2, 0u);
        auto wval = [&](int k, int col) -> float {
          if (k >= 9 * KC2) return 0.0f;
          const int tap = k / KC2, ci = k - tap * KC2;
          const int q = col >> 3, tp = (col & 7) >> 1, u = col & 1;
          if (tp >= R0) return 0.0f;
          const int gate = 2 * q + u;
          return wg_host[0][((size_t)tap * cin + R0 + ci) * 4 * R0 + gate * R0 + tp];
        };
        auto pack = [&](float lo, float hi) -> uint32_t {
          const __half2 v = __floats2half2_rn(lo, hi);
          uint32_t u;
          memcpy(&u, &v, 4);
          return u;
        };
        for (int lane = 0; lane < 32; lane++)
          for (int s_ = 0; s_ < L0_KSTEPS_MAX; s_++)
            for (int q = 0; q < 2; q++) {
              const int gq = lane >> 2, tq = lane & 3, k = 16 * s_ + 2 * tq, col = 8 * q + gq;
              uint32_t *dst = &bf[((size_t)(lane * L0_KSTEPS_MAX + s_) * 2 + q) * 2];
              dst[0] = pack(wval(k, col), wval(k + 1, col));
              dst[1] = pack(wval(k + 8, col), wval(k + 9, col));
            }
        T->l0_bfrag = (uint32_t *)dev_alloc(h, bf.size() * sizeof(uint32_t));
        if (!T->l0_bfrag) return TZ_ENOMEM;
        TZ_CHECK_CUDA(cudaMemcpy(T->l0_bfrag, bf.data(), bf.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        // BM_0 [pixel][gate][channel] -> [pixel][channel][gate]
        const size_t npix = (size_t)h->H[0] * h->W[0];
        std::vector<float> bsrc(npix * 4 * R0), bdst(npix * 4 * R0);
        TZ_CHECK_CUDA(cudaMemcpy(bsrc.data(), h->BM[0], bsrc.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t px = 0; px < npix; px++)
          for (int g = 0; g < 4; g++)
            for (int ch = 0; ch < R0; ch++) bdst[(px * R0 + ch) * 4 + g] = bsrc[(px * 4 + g) * R0 + ch];
        T->l0_bm = (float *)dev_alloc(h, bdst.size() * sizeof(float));
        if (!T->l0_bm) return TZ_ENOMEM;
        TZ_CHECK_CUDA(cudaMemcpy(T->l0_bm, bdst.data(), bdst.size() * sizeof(float), cudaMemcpyHostToDevice));
      }
    }
    if (l < L - 1) {
      // a conv: reads channels [0, 2S_l) of X_l
      std::vector<float> wa((size_t)9 * 2 * h->S[l] * h->S[l + 1]);
      TZ_CHECK_CUDA(cudaMemcpy(wa.data(), h->w_a[l], wa.size() * sizeof(float), cudaMemcpyDeviceToHost));
      std::vector<int> amap(2 * h->S[l]);
      for (int i = 0; i < 2 * h->S[l]; i++) amap[i] = i;
      rc = make_conv(h, &T->aconv[l], 0, l, T->X[l], T->cx[l], amap, wa, 2 * h->S[l], h->S[l + 1], h->S[l + 1]);
      if (rc) return rc;
      ConvArgs &Aa = T->aconv[l].args;
      Aa.bias = h->b_a[l];
      Aa.ahat_next = h->Ahat0[l + 1];
      Aa.xe_out = T->X[l + 1];
      Aa.xe_cstride = T->cx[l + 1];
      Aa.S_next = h->S[l + 1];
      if (tail0) {   // the same convolution reading the other X_0 buffer
        rc = make_conv(h, &T->aconv0_alt, 0, l, T->X0b, T->cx[l], amap, wa, 2 * h->S[l], h->S[l + 1], h->S[l + 1]);
        if (rc) return rc;
        ConvArgs &Ab = T->aconv0_alt.args;
        Ab.bias = Aa.bias;
        Ab.ahat_next = Aa.ahat_next;
        Ab.xe_out = Aa.xe_out;
        Ab.xe_cstride = Aa.xe_cstride;
        Ab.S_next = Aa.S_next;
      }
    }
  }
  return TZ_OK;
}

bool tc_layer0_split(tz_prednet *h) { return h->tc && h->tc->use_gr; }

void tc_destroy(tz_prednet *h) {
  delete h->tc;
  h->tc = nullptr;
}

static int launch_conv(TcState *T, ConvTc *c, int B, cudaStream_t st) {
  ConvArgs A = c->args;
  A.B = B;
  static long long *dbg_buf = nullptr;
  static const bool dbg_on = getenv("TZ_CONV_DEBUG") != nullptr;
  if (dbg_on && !dbg_buf) cudaMalloc(&dbg_buf, 256 * 8 * sizeof(long long));
  A.dbg = dbg_on ? dbg_buf : nullptr;
  const int tiles_b = (B + (1 << A.tb_log) - 1) >> A.tb_log;
  long long n_tiles = (long long)A.tiles_w * A.tiles_h * tiles_b * A.n_tiles_n;
  int grid = n_tiles < T->sm_count ? (int)n_tiles : T->sm_count;
  if (A.halo == 1) grid = grid / A.n_tiles_n * A.n_tiles_n;   // a CTA keeps one N tile: its weights stay in shared memory
  if (A.halo == 4) {
    // clusters of two CTAs; one (pair of M tiles, N tile) unit per cluster and iteration
    long long m_tiles = (long long)A.tiles_w * A.tiles_h * tiles_b;
    n_tiles = ((m_tiles + 1) / 2) * A.n_tiles_n;
    int clusters = T->sm_count / 2;
    if (n_tiles < clusters) clusters = (int)n_tiles;
    grid = 2 * clusters;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128 + 32 * A.epi_warps);
    cfg.dynamicSmemBytes = c->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = (c->epi == 0) ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, true>, c->tmA, c->tmB, A)
                                  : cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, true>, c->tmA, c->tmB, A);
    if (e != cudaSuccess) {
      set_error("cluster launch failed: %s", cudaGetErrorString(e));
      return TZ_ECUDA;
    }
  } else if (c->epi == 0) {
    conv_tc_kernel<0, false><<<grid, 128 + 32 * A.epi_warps, c->smem_bytes, st>>>(c->tmA, c->tmB, A);
  } else if (c->epi == 2) {
    conv_tc_kernel<2, false><<<grid, 128 + 32 * A.epi_warps, c->smem_bytes, st>>>(c->tmA, c->tmB, A);
  } else {
    conv_tc_kernel<1, false><<<grid, 128 + 32 * A.epi_warps, c->smem_bytes, st>>>(c->tmA, c->tmB, A);
  }
  TZ_CHECK_LAUNCH();
  if (dbg_on) {   // diagnostics only: synchronous
    long long hbuf[256 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
    const int issuers = A.halo == 4 ? grid / 2 : grid;
    fprintf(stderr, "[tz conv] epi %d halo %d H %d N %d kchunks %d: mma-thread cycles total %lld, wait tempty %lld, "
            "wait full %lld, wait afull %lld (CTA 0; units per issuing CTA ~%lld)\n", c->epi, A.halo, A.H, A.n_tile,
            A.kchunks, hbuf[0], hbuf[1], hbuf[2], hbuf[3], (n_tiles + issuers - 1) / issuers);
#if TZ_EPI_DEBUG
    fprintf(stderr, "[tz conv]     epilogue warp 4: %lld tiles, cycles waiting for an accumulator %lld, working %lld "
            "(%lld per tile; LSTM: TMEM loads %lld, bias-map loads + gate math %lld); %d epilogue warps in %d group(s), "
            "%d accumulator stages, %d MMA issuer(s)\n", hbuf[6], hbuf[4], hbuf[5], hbuf[6] ? hbuf[5] / hbuf[6] : 0,
            hbuf[7], c->epi == 1 ? hbuf[3] : 0, A.epi_warps, A.epi_groups, A.acc_stages, A.mma_issuers);
#endif
  }
  return TZ_OK;
}

int tc_next(tz_prednet *h, const float *in, float *out, int B, cudaStream_t st, cudaEvent_t *ev, bool skip_e0) {
  TcState *T = h->tc;
  const int L = h->L;
  int ne = 0;
  h->x0_staged = false;
  __half *x0_now = (T->use_gr && T->x0_cur) ? T->X0b : T->X[0];
  if (ev) cudaEventRecord(ev[ne++], st);
  if (skip_e0) {
    if (ev) cudaEventRecord(ev[ne++], st);
  } else {
    const int C = h->S[0];
    long long total = (long long)B * h->H[0] * h->W[0] * C;
    const long long npix = (long long)B * h->H[0] * h->W[0], fpx = (long long)h->H[0] * h->W[0];
    if (T->use_gr && T->cx[0] == 16 && C == 3)
      e0_px_kernel<3><<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(in, h->Ahat0[0], x0_now, npix, fpx);
    else if (T->use_gr && T->cx[0] == 16 && C == 1)
      e0_px_kernel<1><<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(in, h->Ahat0[0], x0_now, npix, fpx);
    else
      e0_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, h->Ahat0[0], x0_now, total,
                                                                   (long long)h->H[0] * h->W[0] * C, C, T->cx[0]);
    TZ_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[ne++], st);
  }
  for (int l = 0; l < L - 1; l++) {
    int rc = launch_conv(T, (l == 0 && T->use_gr && T->x0_cur) ? &T->aconv0_alt : &T->aconv[l], B, st);
    if (rc) return rc;
    if (ev) cudaEventRecord(ev[ne++], st);
  }
  for (int l = L - 1; l >= 0; l--) {
    int rc = launch_conv(T, (l == 0 && T->use_gr) ? &T->rconv0 : &T->gconv[l], B, st);   // layer 0: r_1 -> G
    if (rc) return rc;
    if (ev) cudaEventRecord(ev[ne++], st);
  }
  int rc = TZ_OK;
  if (T->use_gr) {
    // the rest of layer 0 in one CUDA-core kernel: e_0 half of the gates + G + LSTM + A-hat_0 + next e_0
    TZ_REQUIRE(B <= 65535, "tensor-core path: at most 65535 frames per step");
    __half *x0_next = T->x0_cur ? T->X[0] : T->X0b;
    dim3 grid((h->W[0] + 31) / 32, (h->H[0] + 15) / 16, B);
    if (h->S[0] == 3) {
      L0TailW<3> wb;
      memcpy(wb.wa, T->ahat0_w, sizeof(wb.wa));
      memcpy(wb.ba, T->ahat0_b, sizeof(wb.ba));
      l0_tail_kernel<3><<<grid, 320, sizeof(L0TailSmem<3>), st>>>(x0_now, T->cx[0], T->l0_bm, h->C0[0], T->gr, T->l0_bfrag, wb, out, h->H[0],
                                              h->W[0], h->cfg.pixel_max, h->Ahat0[0], x0_next);
    } else {
      L0TailW<1> wb;
      memcpy(wb.wa, T->ahat0_w, sizeof(wb.wa));
      memcpy(wb.ba, T->ahat0_b, sizeof(wb.ba));
      l0_tail_kernel<1><<<grid, 320, sizeof(L0TailSmem<1>), st>>>(x0_now, T->cx[0], T->l0_bm, h->C0[0], T->gr, T->l0_bfrag, wb, out, h->H[0],
                                              h->W[0], h->cfg.pixel_max, h->Ahat0[0], x0_next);
    }
    TZ_CHECK_LAUNCH();
    T->x0_cur ^= 1;
    h->x0_staged = true;
  } else if (h->R[0] == h->S[0] && (h->S[0] == 3 || h->S[0] == 1) && B <= 65535) {
    dim3 grid((h->W[0] + 31) / 32, (h->H[0] + 15) / 16, B);
    if (h->S[0] == 3) {
      Ahat0W<3> wb;
      memcpy(wb.w, T->ahat0_w, sizeof(wb.w));
      memcpy(wb.b, T->ahat0_b, sizeof(wb.b));
      ahat0_kernel<3><<<grid, 256, 0, st>>>(T->r0, wb, out, h->H[0], h->W[0], h->cfg.pixel_max, h->Ahat0[0], T->X[0],
                                            T->cx[0], T->epad[0]);
    } else {
      Ahat0W<1> wb;
      memcpy(wb.w, T->ahat0_w, sizeof(wb.w));
      memcpy(wb.b, T->ahat0_b, sizeof(wb.b));
      ahat0_kernel<1><<<grid, 256, 0, st>>>(T->r0, wb, out, h->H[0], h->W[0], h->cfg.pixel_max, h->Ahat0[0], T->X[0],
                                            T->cx[0], T->epad[0]);
    }
    TZ_CHECK_LAUNCH();
    h->x0_staged = true;
  } else {
    ConvSrc s = {T->r0, h->R[0], 0, 0, (long long)h->H[0] * h->W[0] * h->R[0]};
    rc = conv3x3_direct(&s, 1, h->w_ahat[0], h->R[0], h->S[0], h->b_ahat[0], nullptr, out, B, h->H[0], h->W[0], 2,
                        h->cfg.pixel_max, st);
  }
  if (ev) cudaEventRecord(ev[ne++], st);
  return rc;
}

}  // namespace tz
