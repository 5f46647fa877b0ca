// tezip_b200 -- codec kernels: residual, error-bound quantisation, 1-D delta + histogram, rank map,
// decoder (rank -> symbol -> prefix scan -> reconstruct), DWP metric, key plane.
// HBM-bound integer/byte work: coalesced 16-byte accesses, shared-memory histograms/LUTs, grids sized in
// multiples of the SM count.  Reference lines cited per kernel (paths under /root/reference/src).
#include "tz_codec.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ residual
// compress.py:293-314.  One thread per 8 samples.
template <bool FAST>
__global__ void __launch_bounds__(256) residual_kernel(const uint8_t *__restrict__ frames,
                                                       const float *__restrict__ pool,
                                                       const int32_t *__restrict__ slot, int16_t *__restrict__ x,
                                                       long long n, Geo g) {
  long long ngroups = (n + 7) / 8;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups;
       gi += (long long)gridDim.x * blockDim.x) {
    int v[8];
    resid8<FAST>(frames, pool, slot, g, gi * 8, n, v);
    store8_i16(x, gi * 8, n, v);
  }
}

// ------------------------------------------------------------------------------------------------ error_bound
// compress.py:23-70: greedy interval-intersection scan.  The reference walks one plane serially: it keeps the
// running intersection [l, u] of the intervals [d_i - E_i, d_i + E_i]; when element i would empty it
// (min(u, Du_i) - max(l, Dl_i) < 0) the open segment [head, i) is flushed with trunc((u+l)/2) and a new segment
// starts at i.  Here ONE WARP owns a plane and consumes it 32 elements at a time:
//   A. an inclusive prefix min/max (shuffles), seeded with the carried (u, l), finds where the open segment ends
//      inside the chunk (or absorbs the whole chunk);
//   B. every lane s finds, for a segment that would start at s, the first lane that breaks it (forward walk
//      with shuffles; segments are short, and the walk stops as soon as all lanes are done);
//   C. lane 0's view chases those links from the first break, closing segments until one stays open.
// Identical arithmetic to the reference: IEEE double with explicit _rn intrinsics (nothing is contracted into an
// FMA; NumPy rounds E, d+E, d-E and (u+l)/2 separately).  VAL = int is the exact shortcut for a plane-wide bound
// (abs / rel / absrel): with d integer, min_i fl(d_i+E) = fl(dmin+E) and max_i fl(d_i-E) = fl(dmax-E), so the
// emptiness test depends only on the integer gap dmax - dmin; it equals gap > floor(2E) whenever the additions
// are exact (E a multiple of 2^-36 below 4096) or 2E is at least 1e-6 from an integer (rounding is ~1e-13).
__global__ void __launch_bounds__(128) error_bound_kernel(const uint8_t *__restrict__ frames,
                                                          int16_t *__restrict__ x,
                                                          const uint8_t *__restrict__ apply, long long nt, Geo g,
                                                          int mode, double b0, double b1) {
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per (frame, channel)
  if (t >= nt * g.C) return;
  const long long f = t / g.C;
  const int c = (int)(t - f * g.C);
  if (!apply[f]) return;
  const int lane = threadIdx.x & 31;
  const int n = g.H * g.W;
  const int C = g.C;
  const uint8_t *o = frames + f * g.frame_elems + c;
  int16_t *d = x + f * g.frame_elems + c;
  double E = 0.0;
  if (mode == TZ_MODE_ABS) {
    E = fabs(b0);                                                        // :29
  } else if (mode == TZ_MODE_REL || mode == TZ_MODE_ABSREL) {
    int mx = 0, mn = 255;                                                // :31-32 / :36-37
    for (int i = lane; i < n; i += 32) {
      int v = o[(long long)i * C];
      mx = max(mx, v);
      mn = min(mn, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    }
    if (mode == TZ_MODE_REL) {
      E = __dmul_rn((double)(mx - mn), b0);                              // :33
    } else {
      double a = fabs(b0), r = __dmul_rn((double)(mx - mn), b1);         // :38-39
      E = (a < r) ? a : r;                                               // :40-43
    }
  }
  if (mode != TZ_MODE_PWREL) {
    const double twoE = E + E;
    const double sc = E * 68719476736.0;   // 2^36
    const bool exact = (E < 4096.0) && (sc == floor(sc));
    const bool clear = fabs(twoE - rint(twoE)) > 1e-6;
    if (E >= 0.0 && (exact || clear)) {
      const int G = twoE >= 70000.0 ? 70000 : (int)floor(twoE);
      eb_plane_warp<EbInt>(o, d, n, C, false, b0, E, G);
      return;
    }
  }
  eb_plane_warp<EbDbl>(o, d, n, C, mode == TZ_MODE_PWREL, b0, E, 0);
}

// ---- tiled variant for plane-wide bounds (abs / rel / absrel): one CTA per plane ---------------------------------
// The warp-serial kernel above is bound by the latency of its dependent chain (one chunk after another, one
// segment after another inside a chunk); with short segments (noisy residuals) that chain is ~7k cycles per 32
// elements.  Here everything that does not depend on where the previous segment ended is done in parallel, for
// EVERY element as if a segment started there, and the serial part shrinks to one ballot per 32-element chunk:
//   S1 (all warps, one chunk of 32 elements per warp pass)
//        pre[j]   inclusive in-chunk prefix state (min d, max d)                       -- 5 shuffle rounds
//        nb[j]    first in-chunk element that breaks a segment started at j, or 32      -- sparse table + binary lifting
//        run[j]   state of [j, nb[j])                                                   --   (5 + 5 shuffles)
//        last[j]  last start inside the chunk of the chain j -> nb[j] -> nb[nb[j]] ...  -- 5 pointer-jumping rounds
//   S2 (warp 0) walks the chunks with the open segment's state X: ballot(broken(join(X, pre[j]))) gives the break
//        b; the orbit then enters the chunk at b, leaves it at last[b] with X = run[last[b]].
//   S3 (all warps) marks the starts inside each chunk from its entry (reachability doubling with a warp OR),
//        picks every element's segment state and writes trunc((u+l)/2).
// A segment still open at the end of a tile is written when it closes (or at the end of the plane).
// State is packed as two int16: lo = min d, hi = ~(max d), so that join is ONE __vmins2 (VIMNMX.S16X2) and
// (min - max + G + 1) is ONE __dp2a_lo; ~ maps int16 onto int16, so every int16 residual is representable.
// scripts/eb_tile_emulate.py is a lane-level numpy emulation of this kernel checked against the oracle.
// EB_T elements per tile (EB_NCH = EB_T / 32 chunks, at most 32), EB_THREADS threads per CTA: template parameters,
// the launcher instantiates a few shapes (TZ_EB_CFG selects one for experiments).
constexpr uint32_t EBP_ID = 0x7fff7fffu;   // identity of join: (min = 32767, max = -32768)

__device__ __forceinline__ uint32_t ebp_pack(int d) { return ((uint32_t)d & 0xffffu) | ((uint32_t)(~d) << 16); }
__device__ __forceinline__ uint32_t ebp_join(uint32_t p, uint32_t q) { return __vmins2(p, q); }
// max - min > G  <=>  min + ~max + 1 + G < 0
__device__ __forceinline__ bool ebp_broken(uint32_t p, int G1) { return __dp2a_lo((int)p, 0x0101, G1) < 0; }
// compress.py:61 as EbInt::mid.  EXACT: E is a multiple of 2^-36 below 4096, so fl(a+E), fl(b-E) and their sum are
// exact and trunc(((a+E)+(b-E))/2) is the truncating integer division (a+b)/2.
template <bool EXACT>
__device__ __forceinline__ int16_t ebp_mid(uint32_t p, double E) {
  const int a = (int)(int16_t)(p & 0xffffu);
  const int b = ~(int)(int16_t)(p >> 16);
  if (EXACT) return (int16_t)((a + b) / 2);
  return (int16_t)(long long)__dmul_rn(__dadd_rn(__dadd_rn((double)a, E), __dsub_rn((double)b, E)), 0.5);
}

template <int EB_T>
struct EbTileSmem {
  static constexpr int EB_NCH = EB_T / 32;
  uint32_t pre[EB_T];      // S1 -> S2: inclusive in-chunk prefix state
  uint32_t run[EB_T];      // S1 -> S3: state of the in-chunk segment [j, nb[j])
  uint32_t exitst[EB_T];   // S1 -> S2: run[last[j]], the state with which the orbit of j leaves the chunk
  uint16_t meta[EB_T];     // lo 8: nb (1..32), hi 8: last
  uint32_t inst[EB_NCH];   // S2 -> S3: state of the segment covering the elements before the chunk's entry
  uint32_t outst[EB_NCH];  // S2 -> S3: state of the segment that starts at the chunk's last start
  uint8_t entry[EB_NCH];   // S2 -> S3: first start inside the chunk (0xff: none)
  uint32_t inx[EB_NCH];    // S2 -> S2b: state of the open segment on entry to a chunk that closes it
  int8_t prevc[EB_NCH];    // S2 -> S2b: chunk of the previous close in this tile (-1: none)
  int red[2 * 8];
  int head, open_from, wb_head;
  uint32_t carry, wb_state;
};

template <bool EXACT, int EB_T, int EB_THREADS>
__device__ __forceinline__ void eb_plane_tiles(EbTileSmem<EB_T> &sm, int16_t *__restrict__ d, int n, int C, double E,
                                               int G1, int warp, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int EB_NCH = EB_T / 32, EB_WARPS = EB_THREADS / 32;
  const int tid = warp * 32 + lane;
  if (tid == 0) {
    sm.head = 0;
    sm.carry = EBP_ID;
  }
  for (int tile = 0; tile < n; tile += EB_T) {
    const int Tn = min(EB_T, n - tile);
    const int nch = (Tn + 31) >> 5;
    // ---------------------------------------------------------------- S1
    {
      // two chunks per pass (c and c + EB_WARPS): their shuffle chains are independent, which doubles the
      // instruction-level parallelism of a warp
      constexpr int U = 2;
      const int16_t *dp = d + (long long)(tile + tid) * C;
      for (int c = warp; c < nch; c += U * EB_WARPS, dp += (long long)U * EB_THREADS * C) {
        int i[U];
        uint32_t v[U], pre[U], M[5][U], run[U], pk[U];
        int pos[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          i[u] = 32 * (c + u * EB_WARPS) + lane;
          v[u] = (i[u] < Tn) ? ebp_pack((int)dp[(long long)u * EB_THREADS * C]) : EBP_ID;
          pre[u] = v[u];
          M[0][u] = v[u];
          run[u] = v[u];
          pos[u] = lane + 1;
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
#pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t t = __shfl_up_sync(FULL, pre[u], off);
            pre[u] = ebp_join(pre[u], (lane >= off) ? t : EBP_ID);
          }
#pragma unroll
        for (int k = 0; k < 4; k++)   // M[k][j] = join(v[j .. j + 2^k)), identity beyond the chunk
#pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t t = __shfl_down_sync(FULL, M[k][u], 1 << k);
            M[k + 1][u] = ebp_join(M[k][u], (lane + (1 << k) < 32) ? t : EBP_ID);
          }
#pragma unroll
        for (int k = 4; k >= 0; k--)   // binary lifting: the longest unbroken [lane, pos)
#pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t nx = ebp_join(run[u], __shfl_sync(FULL, M[k][u], pos[u]));
            const bool ok = (pos[u] + (1 << k) <= 32) & !ebp_broken(nx, G1);
            run[u] = ok ? nx : run[u];
            pos[u] = ok ? pos[u] + (1 << k) : pos[u];
          }
#pragma unroll
        for (int u = 0; u < U; u++) pk[u] = ((uint32_t)pos[u] << 8) | (uint32_t)lane;   // (next start, last start visited)
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t t = __shfl_sync(FULL, pk[u], pk[u] >> 8);
            pk[u] = (pk[u] < (32u << 8)) ? t : pk[u];
          }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t ex = __shfl_sync(FULL, run[u], pk[u]);   // lane = pk & 31 = last
          if (c + u * EB_WARPS < nch) {   // warp-uniform
            sm.pre[i[u]] = pre[u];
            sm.run[i[u]] = run[u];
            sm.exitst[i[u]] = ex;
            sm.meta[i[u]] = (uint16_t)((uint32_t)pos[u] | ((pk[u] & 0xffu) << 8));
            if (lane == 0) sm.entry[c + u * EB_WARPS] = 0xff;
          }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- S2
    if (warp == 0) {
      // The serial chain carries only what the next hop needs: X (state of the open segment) and, per chunk that
      // closes a segment, three one-word records (entry lane, X on entry, chunk of the previous close).  The states
      // with which segments close -- and which chunks they cover -- are derived from those records afterwards, by
      // the 32 lanes in parallel (S2b).
      uint32_t X = sm.carry;
      const int hd0 = sm.head;
      int seg_chunk = -1;   // chunk of this tile in which the open segment starts (-1: before the tile)
      int seg_b = 0;        // its entry lane
      uint32_t pj = sm.pre[lane];
      for (int c = 0; c < nch; c++) {
        const uint32_t tj = ebp_join(X, pj);
        if (c + 1 < nch) pj = sm.pre[32 * (c + 1) + lane];   // does not depend on the chain: overlaps the ballot
        const unsigned m = __ballot_sync(FULL, ebp_broken(tj, G1));
        if (m == 0) {   // the whole chunk joins the open segment
          X = __shfl_sync(FULL, tj, 31);
          continue;
        }
        const int b = __ffs(m) - 1;
        const uint32_t nextX = sm.exitst[32 * c + b];
        if (lane == 0) {
          sm.entry[c] = (uint8_t)b;
          sm.inx[c] = X;
          sm.prevc[c] = (int8_t)seg_chunk;
        }
        X = nextX;
        seg_chunk = c;
        seg_b = b;
      }
      if (lane == 0) {
        int hd = hd0;
        if (seg_chunk >= 0) hd = tile + 32 * seg_chunk + (sm.meta[32 * seg_chunk + seg_b] >> 8);
        sm.carry = X;
        sm.head = hd;
        sm.open_from = (hd >= tile) ? hd - tile : 0;
        sm.wb_head = -1;
      }
      __syncwarp();
      // ---- S2b: lane c owns chunk c.  A chunk with an entry b closes the segment that was open on entry: its state
      // is X_in joined with the chunk's prefix up to b-1; it covers the chunks since the previous close.
      static_assert(EB_NCH <= 32, "one lane per chunk");
      if (lane < nch) {
        const int c = lane;
        const int b = sm.entry[c];
        if (b != 0xff) {
          const uint32_t xin = sm.inx[c];
          const uint32_t Xc = (b > 0) ? ebp_join(xin, sm.pre[32 * c + b - 1]) : xin;
          const int pc = sm.prevc[c];
          sm.inst[c] = Xc;
          if (pc >= 0) {
            sm.outst[pc] = Xc;
          } else {   // the segment began before this tile (or is the first of the plane)
            sm.wb_head = hd0;
            sm.wb_state = Xc;
          }
          for (int cc = pc + 1; cc < c; cc++) sm.inst[cc] = Xc;   // chunks absorbed whole
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- S3
    const int open_from = sm.open_from;
    {
      const int wbh = sm.wb_head;
      if (wbh >= 0 && wbh < tile) {   // the part of the first closed segment that lies in earlier tiles
        const int16_t q = ebp_mid<EXACT>(sm.wb_state, E);
        for (int j = wbh + tid; j < tile; j += EB_THREADS) d[(long long)j * C] = q;
      }
    }
    {
      constexpr int U = 2;   // two chunks per pass, as in S1
      int16_t *dp = d + (long long)(tile + tid) * C;
      for (int c = warp; c < nch && 32 * c < open_from; c += U * EB_WARPS, dp += (long long)U * EB_THREADS * C) {
        int i[U], ent[U], nb[U], J[U];
        uint32_t st[U], run[U];
        unsigned R[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int cu = c + u * EB_WARPS;
          live[u] = cu < nch && 32 * cu < open_from;   // warp-uniform
          const int cc = live[u] ? cu : c;             // a dead second chunk replays the first (results unused)
          i[u] = 32 * cc + lane;
          ent[u] = sm.entry[cc];
          st[u] = sm.inst[cc];
          nb[u] = sm.meta[i[u]] & 0xff;
          run[u] = sm.run[i[u]];
          R[u] = (ent[u] != 0xff) ? 1u << ent[u] : 0u;   // starts reachable from the entry
          J[u] = nb[u];
        }
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
          for (int u = 0; u < U; u++) {
            const unsigned tgt = ((R[u] >> lane) & 1u) << (J[u] & 31);
            R[u] |= __reduce_or_sync(FULL, (J[u] < 32) ? tgt : 0u);
            const int t = __shfl_sync(FULL, J[u], J[u]);
            J[u] = (J[u] < 32) ? t : 32;
          }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int cu = live[u] ? c + u * EB_WARPS : c;
          const int s = 31 - __clz((int)(R[u] & (0xffffffffu >> (31 - lane))));   // -1 for lanes before the entry
          const int nb_s = __shfl_sync(FULL, nb[u], s);
          const uint32_t run_s = __shfl_sync(FULL, run[u], s);
          const uint32_t own = (nb_s >= 32) ? sm.outst[cu] : run_s;
          if (ent[u] != 0xff && lane >= ent[u]) st[u] = own;
          if (live[u] && i[u] < min(Tn, open_from)) dp[(long long)u * EB_THREADS * C] = ebp_mid<EXACT>(st[u], E);
        }
      }
    }
    __syncthreads();
  }
  {   // compress.py:67: the segment still open at the end of the plane
    const int head = sm.head;
    const int16_t q = ebp_mid<EXACT>(sm.carry, E);
    for (int j = head + tid; j < n; j += EB_THREADS) d[(long long)j * C] = q;
  }
}

template <int EB_T, int EB_THREADS>
__global__ void __launch_bounds__(EB_THREADS) error_bound_tiles_kernel(const uint8_t *__restrict__ frames,
                                                                       int16_t *__restrict__ x,
                                                                       const uint8_t *__restrict__ apply,
                                                                       long long nt, Geo g, int mode, double b0,
                                                                       double b1) {
  __shared__ EbTileSmem<EB_T> sm;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int EB_WARPS = EB_THREADS / 32;
  const long long t = blockIdx.x;        // one CTA per (frame, channel)
  const long long f = t / g.C;
  const int ch = (int)(t - f * g.C);
  if (!apply[f]) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULL, tid >> 5, 0);   // warp-uniform for the compiler: no divergence guards on shuffles
  const int n = g.H * g.W;
  const int C = g.C;
  const uint8_t *o = frames + f * g.frame_elems + ch;
  int16_t *d = x + f * g.frame_elems + ch;
  double E = fabs(b0);                                                   // :29
  if (mode != TZ_MODE_ABS) {
    int mx = 0, mn = 255;                                                // :31-32 / :36-37
    for (int i = tid; i < n; i += EB_THREADS) {
      int v = o[(long long)i * C];
      mx = max(mx, v);
      mn = min(mn, v);
    }
    mx = __reduce_max_sync(FULL, mx);
    mn = __reduce_min_sync(FULL, mn);
    if (lane == 0) {
      sm.red[2 * warp] = mx;
      sm.red[2 * warp + 1] = mn;
    }
    __syncthreads();
    for (int w = 0; w < EB_WARPS; w++) {
      mx = max(mx, sm.red[2 * w]);
      mn = min(mn, sm.red[2 * w + 1]);
    }
    if (mode == TZ_MODE_REL) {
      E = __dmul_rn((double)(mx - mn), b0);                              // :33
    } else {
      double a = fabs(b0), r = __dmul_rn((double)(mx - mn), b1);         // :38-39
      E = (a < r) ? a : r;                                               // :40-43
    }
  }
  const double twoE = E + E;
  const double sc = E * 68719476736.0;   // 2^36
  const bool exact = (E < 4096.0) && (sc == floor(sc));
  const bool clear = fabs(twoE - rint(twoE)) > 1e-6;
  if (!(E >= 0.0 && (exact || clear))) {   // the integer shortcut does not apply: IEEE-double scan, one warp
    if (warp == 0) eb_plane_warp<EbDbl>(o, d, n, C, false, b0, E, 0);
    return;
  }
  const int G1 = (twoE >= 70000.0 ? 70000 : (int)floor(twoE)) + 1;
  if (exact)
    eb_plane_tiles<true, EB_T, EB_THREADS>(sm, d, n, C, E, G1, warp, lane);
  else
    eb_plane_tiles<false, EB_T, EB_THREADS>(sm, d, n, C, E, G1, warp, lane);
}

// ------------------------------------------------------------------------------------------------ table on the device
// compress.py:352-361 + :84-90 without the host round trip between the histogram and the rank-map pass: one CTA
// compacts the non-empty bins, ranks every one of them by counting the bins that sort before it (count descending,
// ties by ascending symbol -- keys are distinct, so the rank is the table position), and writes the symbol -> rank
// LUT.  The LUT is the plain scatter only when no symbol lies inside the rank range [0, n) (the reference's
// sequential where() passes chain otherwise, ops._sequential_replace): meta[1] reports that case and the host redoes
// the LUT.  meta[0] = table length.
__global__ void __launch_bounds__(1024) build_table_kernel(const unsigned long long *__restrict__ hist,
                                                           int16_t *__restrict__ table, int16_t *__restrict__ lut,
                                                           int32_t *__restrict__ meta) {
  __shared__ unsigned long long keys[TZ_HIST_BINS];   // count * 4096 + (4095 - symbol): larger sorts first
  __shared__ int n_s, bad_s;
  if (threadIdx.x == 0) {
    n_s = 0;
    bad_s = 0;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < TZ_HIST_BINS; s += blockDim.x) {
    lut[s] = (int16_t)s;
    const unsigned long long c = hist[s];
    if (c) keys[atomicAdd(&n_s, 1)] = c * TZ_HIST_BINS + (unsigned long long)(TZ_HIST_BINS - 1 - s);
  }
  __syncthreads();
  const int n = n_s;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long k = keys[i];
    int r = 0;
    for (int j = 0; j < n; j++) r += keys[j] > k;
    const int sym = TZ_HIST_BINS - 1 - (int)(k % TZ_HIST_BINS);
    table[r] = (int16_t)sym;
    if (sym < n) bad_s = 1;   // a symbol inside the rank range: the scatter is not the reference's LUT
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) lut[table[i]] = (int16_t)i;   // after the identity fill above
  if (threadIdx.x == 0) {
    meta[0] = n;
    meta[1] = bad_s;
  }
}

// ------------------------------------------------------------------------------------------------ stream kernels
// Work decomposition shared by the two encoder passes and the decoder: one WARP owns a chunk of WCH = 1024
// consecutive stream elements and walks it in four coalesced rounds of 256 (lane k: 8 consecutive elements = one
// 16-byte access of the int16 stream).  All loads of the four rounds are issued before the first use (the
// previous one-group-per-thread loops kept 16 bytes per thread in flight, half of what the HBM latency needs);
// the element before a lane's group comes from the neighbouring lane by shuffle -- only the first element of a
// chunk is fetched again -- and the (frame, offset) pair of a group is derived from ONE 64-bit division per chunk
// (it used to be two per group: 64-bit integer division is ~100 instructions and made these passes issue-bound).
// No block-level synchronisation inside the loop.
constexpr int WCH_ROUNDS = 4;
constexpr int WCH = 32 * 8 * WCH_ROUNDS;

struct StreamPos {   // stream element i = f * frame_elems + r
  long long f;
  unsigned int r;
};
__device__ __forceinline__ StreamPos stream_locate(long long i, const Geo &g) {
  StreamPos p;
  p.f = i / g.frame_elems;
  p.r = (unsigned int)rem_in_frame(i, p.f, g.frame_elems);
  return p;
}
__device__ __forceinline__ StreamPos stream_advance(StreamPos p, unsigned int d, unsigned int FE) {   // r + d < 2^32
  p.r += d;
  if (p.r >= FE) {
    const unsigned int q = p.r / FE;
    p.f += q;
    p.r -= q * FE;
  }
  return p;
}
// offset of in-frame element r inside the (padded) prediction frame
__device__ __forceinline__ unsigned int pool_ofs(unsigned int r, const Geo &g) {
  if (g.prow == g.rowlen) return r;
  const unsigned int row = r / (unsigned int)g.rowlen;
  return row * (unsigned int)g.prow + (r - row * (unsigned int)g.rowlen);
}
// compress.py:293-314 for one sample at a known position.
__device__ __forceinline__ int resid_pos(const uint8_t *__restrict__ frames, const float *__restrict__ pool,
                                         const int32_t *__restrict__ slot, const Geo &g, StreamPos p) {
  const int s = slot[p.f];
  if (s < 0) return 0;
  const float pv = pool[(long long)s * g.pframe_elems + pool_ofs(p.r, g)];
  return q255(pv) - (int)frames[p.f * g.frame_elems + p.r];
}
__device__ __forceinline__ void unpack8_i16(const uint4 a, int v[8]) {
  v[0] = (int16_t)(a.x & 0xffff); v[1] = (int16_t)(a.x >> 16);
  v[2] = (int16_t)(a.y & 0xffff); v[3] = (int16_t)(a.y >> 16);
  v[4] = (int16_t)(a.z & 0xffff); v[5] = (int16_t)(a.z >> 16);
  v[6] = (int16_t)(a.w & 0xffff); v[7] = (int16_t)(a.w >> 16);
}
__device__ __forceinline__ uint4 pack8_i16(const int v[8]) {
  uint4 a;
  a.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
  a.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
  a.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
  a.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
  return a;
}
__device__ __forceinline__ void unpack8_u8(const uint2 a, int v[8]) {
  v[0] = a.x & 0xff; v[1] = (a.x >> 8) & 0xff; v[2] = (a.x >> 16) & 0xff; v[3] = a.x >> 24;
  v[4] = a.y & 0xff; v[5] = (a.y >> 8) & 0xff; v[6] = (a.y >> 16) & 0xff; v[7] = a.y >> 24;
}

// The quantised residuals of one chunk: v[round][k] = x[cbase + round * 256 + lane * 8 + k] (0 beyond n), and, in
// lane 0, prev0 = the element before the chunk (compress.py:75; on a shard *prev_x; 0 at the start of the stream).
// SRC 0: x is materialised (int16).  SRC 1 / 2: the residual is recomputed from frames + predictions (fused
// lossless path); 2 = rowlen % 8 == 0, so a group of 8 lies inside one row of one frame.
template <int SRC>
__device__ __forceinline__ void enc_chunk_load(const int16_t *__restrict__ x, const uint8_t *__restrict__ frames,
                                               const float *__restrict__ pool, const int32_t *__restrict__ slot,
                                               const Geo &g, long long n, long long cbase, int lane, int has_prev,
                                               const int32_t *__restrict__ prev_x, int v[WCH_ROUNDS][8], int &prev0) {
  prev0 = 0;
  if (SRC == 0) {
    if (cbase + WCH <= n) {
      uint4 q[WCH_ROUNDS];
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) q[it] = *reinterpret_cast<const uint4 *>(x + cbase + it * 256 + lane * 8);
      if (lane == 0) prev0 = (cbase > 0 || has_prev == 2) ? (int)x[cbase - 1] : 0;   // has_prev == 2: x points into a longer stream
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) unpack8_i16(q[it], v[it]);
    } else {
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) load8_i16(x, cbase + it * 256 + lane * 8, n, v[it]);
      if (lane == 0) prev0 = (cbase > 0 || has_prev == 2) ? (int)x[cbase - 1] : 0;
    }
  } else if (SRC == 2) {
    const unsigned int FE = (unsigned int)g.frame_elems;
    const StreamPos p0 = stream_locate(cbase, g);
    uint2 a[WCH_ROUNDS];
    float4 pa[WCH_ROUNDS], pb[WCH_ROUNDS];
    bool on[WCH_ROUNDS];
#pragma unroll
    for (int it = 0; it < WCH_ROUNDS; it++) {
      const unsigned int off = it * 256 + lane * 8;
      const long long i0 = cbase + off;
      on[it] = false;
      if (i0 < n) {   // n is a multiple of 8 here: a group is entirely inside or outside the stream
        const StreamPos p = stream_advance(p0, off, FE);
        const int s = slot[p.f];
        if (s >= 0) {   // s < 0: first frame of a window, x = 0 (compress.py:314)
          on[it] = true;
          a[it] = *reinterpret_cast<const uint2 *>(frames + i0);
          const float4 *pp = reinterpret_cast<const float4 *>(pool + (long long)s * g.pframe_elems + pool_ofs(p.r, g));
          pa[it] = pp[0];
          pb[it] = pp[1];
        }
      }
    }
    if (lane == 0 && cbase > 0) {
      StreamPos pm = p0;
      if (pm.r > 0) {
        pm.r--;
      } else {
        pm.f--;
        pm.r = FE - 1;
      }
      prev0 = resid_pos(frames, pool, slot, g, pm);
    }
#pragma unroll
    for (int it = 0; it < WCH_ROUNDS; it++) {
      if (on[it]) {
        int o[8];
        unpack8_u8(a[it], o);
        v[it][0] = q255(pa[it].x) - o[0]; v[it][1] = q255(pa[it].y) - o[1];
        v[it][2] = q255(pa[it].z) - o[2]; v[it][3] = q255(pa[it].w) - o[3];
        v[it][4] = q255(pb[it].x) - o[4]; v[it][5] = q255(pb[it].y) - o[5];
        v[it][6] = q255(pb[it].z) - o[6]; v[it][7] = q255(pb[it].w) - o[7];
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) v[it][k] = 0;
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < WCH_ROUNDS; it++) resid8<false>(frames, pool, slot, g, cbase + it * 256 + lane * 8, n, v[it]);
    if (lane == 0 && cbase > 0) prev0 = resid_at(frames, pool, slot, g, cbase - 1);
  }
  if (lane == 0 && cbase == 0 && has_prev == 1) prev0 = *prev_x;
}

// y of round `it` (compress.py:75): the element before a lane's group is the neighbouring lane's last element, the
// previous round's last element (lane 0), or prev0 (lane 0, round 0).
__device__ __forceinline__ void enc_round_delta(const int v[WCH_ROUNDS][8], int it, int lane, int prev0,
                                                bool first_global, int y[8]) {
  const int up = __shfl_up_sync(0xffffffffu, v[it][7], 1);
  const int wrap = __shfl_sync(0xffffffffu, it > 0 ? v[it - 1][7] : 0, 31);
  const int prev = lane > 0 ? up : (it > 0 ? wrap : prev0);
  delta8(v[it], prev, first_global && it == 0 && lane == 0, y);
}

// ------------------------------------------------------------------------------------------------ delta + histogram
// compress.py:73-77 + :348-355.  Shared-memory histogram (16 KB), run-length aggregated atomics, one 64-bit global
// atomic per non-empty bin per block.
template <int SRC>
__global__ void __launch_bounds__(256, SRC == 0 ? 3 : 2) delta_hist_kernel(const int16_t *__restrict__ x,
                                                            const uint8_t *__restrict__ frames,
                                                            const float *__restrict__ pool,
                                                            const int32_t *__restrict__ slot, Geo g, long long n,
                                                            int has_prev, const int32_t *__restrict__ prev_x,
                                                            unsigned long long *__restrict__ hist,
                                                            unsigned long long *__restrict__ overflow) {
  __shared__ unsigned int sh[TZ_HIST_BINS];
  __shared__ unsigned int sh_ovf;
  for (int i = threadIdx.x; i < TZ_HIST_BINS; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x == 0) sh_ovf = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nchunks = (n + WCH - 1) / WCH;
  for (long long wid = (long long)blockIdx.x * 8 + warp; wid < nchunks; wid += (long long)gridDim.x * 8) {
    const long long cbase = wid * WCH;
    int v[WCH_ROUNDS][8], prev0;
    enc_chunk_load<SRC>(x, frames, pool, slot, g, n, cbase, lane, has_prev, prev_x, v, prev0);
#pragma unroll
    for (int it = 0; it < WCH_ROUNDS; it++) {
      const long long i0 = cbase + it * 256 + lane * 8;
      int y[8];
      enc_round_delta(v, it, lane, prev0, cbase == 0 && !has_prev, y);
      int cur = -1, cnt = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (i0 + k < n) {
          int s = (int)(int16_t)(TZ_SYMBOL_OFFSET - y[k]);                 // :348 int16 arithmetic
          if (s == cur) {
            cnt++;
          } else {
            if (cnt) {
              if ((unsigned)cur < TZ_HIST_BINS) atomicAdd(&sh[cur], cnt); else atomicAdd(&sh_ovf, cnt);
            }
            cur = s;
            cnt = 1;
          }
        }
      }
      if (cnt) {
        if ((unsigned)cur < TZ_HIST_BINS) atomicAdd(&sh[cur], cnt); else atomicAdd(&sh_ovf, cnt);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TZ_HIST_BINS; i += blockDim.x) {
    unsigned int c = sh[i];
    if (c) atomicAdd(&hist[i], (unsigned long long)c);
  }
  if (threadIdx.x == 0 && sh_ovf) atomicAdd(overflow, (unsigned long long)sh_ovf);
}

// ------------------------------------------------------------------------------------------------ delta + rank map
// compress.py:84-90 with the symbol -> rank table held in shared memory (pure LUT: SURVEY.md A13).
__device__ __forceinline__ void load_lut_smem(int16_t *sl, const int16_t *__restrict__ lut) {   // 8 KB, 16-byte accesses
  for (int i = threadIdx.x; i < TZ_HIST_BINS / 8; i += blockDim.x)
    reinterpret_cast<uint4 *>(sl)[i] = reinterpret_cast<const uint4 *>(lut)[i];
  __syncthreads();
}

template <int SRC>
__global__ void __launch_bounds__(256, SRC == 0 ? 3 : 2) delta_rank_kernel(const int16_t *__restrict__ x,
                                                            const uint8_t *__restrict__ frames,
                                                            const float *__restrict__ pool,
                                                            const int32_t *__restrict__ slot, Geo g, long long n,
                                                            int has_prev, const int32_t *__restrict__ prev_x,
                                                            const int16_t *__restrict__ lut,
                                                            int16_t *__restrict__ out) {
  __shared__ __align__(16) int16_t sl[TZ_HIST_BINS];
  if (lut) load_lut_smem(sl, lut);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nchunks = (n + WCH - 1) / WCH;
  for (long long wid = (long long)blockIdx.x * 8 + warp; wid < nchunks; wid += (long long)gridDim.x * 8) {
    const long long cbase = wid * WCH;
    int v[WCH_ROUNDS][8], prev0;
    enc_chunk_load<SRC>(x, frames, pool, slot, g, n, cbase, lane, has_prev, prev_x, v, prev0);
#pragma unroll
    for (int it = 0; it < WCH_ROUNDS; it++) {
      int y[8];
      enc_round_delta(v, it, lane, prev0, cbase == 0 && !has_prev, y);
      if (lut) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          int s = (int)(int16_t)(TZ_SYMBOL_OFFSET - y[k]);
          y[k] = ((unsigned)s < TZ_HIST_BINS) ? (int)sl[s] : s;            // where() leaves other values alone
        }
      }
      store8_i16(out, cbase + it * 256 + lane * 8, n, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------ decoder

// decompress.py:31-36,236: rank -> symbol (LUT) -> y = 1600 - s; or y = body with -n streams.
__device__ __forceinline__ void map8(const int16_t *sl, bool use_lut, int v[8]) {
  if (use_lut) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      int r = v[k];
      int s = ((unsigned)r < TZ_HIST_BINS) ? (int)sl[r] : r;
      v[k] = (int)(int16_t)(TZ_SYMBOL_OFFSET - s);
    }
  }
}

// pass 1: sums[c] = sum of y over chunk c (uint32 wrap-around), zero for the padding chunks c in [nchunks, npad)
__global__ void __launch_bounds__(256) decode_wsum_kernel(const int16_t *__restrict__ body, long long n, long long npad,
                                                          int use_lut, const int16_t *__restrict__ lut,
                                                          unsigned int *__restrict__ sums) {
  __shared__ __align__(16) int16_t sl[TZ_HIST_BINS];
  if (use_lut) load_lut_smem(sl, lut);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long wid = (long long)blockIdx.x * 8 + warp; wid < npad; wid += (long long)gridDim.x * 8) {
    const long long cbase = wid * WCH;
    unsigned int s = 0;
    if (cbase + WCH <= n) {
      uint4 q[WCH_ROUNDS];
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) q[it] = *reinterpret_cast<const uint4 *>(body + cbase + it * 256 + lane * 8);
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) {
        int v[8];
        unpack8_i16(q[it], v);
        map8(sl, use_lut != 0, v);
#pragma unroll
        for (int k = 0; k < 8; k++) s += (unsigned int)v[k];
      }
    } else {
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) {
        const long long i0 = cbase + it * 256 + lane * 8;
        int v[8];
        load8_i16(body, i0, n, v);   // zeros beyond n
        map8(sl, use_lut != 0, v);
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (i0 + k < n) s += (unsigned int)v[k];
      }
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) sums[wid] = s;
  }
}

// pass 2: exclusive scan of the chunk sums in place, one block; 8 entries per thread (two 16-byte accesses), two
// barriers per 8192 entries.  npad is a multiple of 8.
__global__ void __launch_bounds__(1024) decode_scan8_kernel(unsigned int *__restrict__ sums, long long npad) {
  __shared__ unsigned int wtot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int carry = 0;   // the same value in every thread
  for (long long base = 0; base < npad; base += 8192) {
    const long long i = base + (long long)threadIdx.x * 8;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
    if (i < npad) {
      a = *reinterpret_cast<const uint4 *>(sums + i);
      b = *reinterpret_cast<const uint4 *>(sums + i + 4);
    }
    const unsigned int tot = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    unsigned int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    const unsigned int w = wtot[lane];   // every warp scans the 32 warp totals for itself
    unsigned int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    const unsigned int woff = __shfl_sync(0xffffffffu, winc - w, warp);
    const unsigned int total = __shfl_sync(0xffffffffu, winc, 31);
    unsigned int e = carry + woff + inc - tot;
    if (i < npad) {
      uint4 oa, ob;
      oa.x = e; e += a.x; oa.y = e; e += a.y; oa.z = e; e += a.z; oa.w = e; e += a.w;
      ob.x = e; e += b.x; ob.y = e; e += b.y; ob.z = e; e += b.z; ob.w = e;
      *reinterpret_cast<uint4 *>(sums + i) = oa;
      *reinterpret_cast<uint4 *>(sums + i + 4) = ob;
    }
    carry += total;
    __syncthreads();   // wtot is rewritten in the next round
  }
}

// pass 3: decompress.py:22-29 (as a prefix sum), :252-256, :269.  x[i] = x0 + y0 - S[i], S inclusive from element 0.
template <bool FAST>
__global__ void __launch_bounds__(256, 2) decode_wrecon_kernel(
    const int16_t *__restrict__ body, long long n, long long nchunks, int use_lut, const int16_t *__restrict__ lut,
    const unsigned int *__restrict__ chunk_prefix, int first_mode, int first_x, const float *__restrict__ pool,
    const int32_t *__restrict__ slot, const uint8_t *__restrict__ key_plane, uint8_t *__restrict__ out,
    int16_t *__restrict__ x_out, Geo g) {
  __shared__ __align__(16) int16_t sl[TZ_HIST_BINS];
  if (use_lut) load_lut_smem(sl, lut);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // y[0] of the whole stream (needed because x[i] = x0 + y0 - S[i])
  int y0;
  {
    int r = body[0];
    if (use_lut) {
      int s = ((unsigned)r < TZ_HIST_BINS) ? (int)sl[r] : r;
      y0 = (int)(int16_t)(TZ_SYMBOL_OFFSET - s);
    } else {
      y0 = r;
    }
  }
  const int x0 = (first_mode == 0) ? y0 : first_x;
  const unsigned int xy = (unsigned int)(x0 + y0);
  const unsigned int FE = (unsigned int)g.frame_elems;
  for (long long wid = (long long)blockIdx.x * 8 + warp; wid < nchunks; wid += (long long)gridDim.x * 8) {
    const long long cbase = wid * WCH;
    unsigned int before = chunk_prefix[wid];   // sum of all y before the current round
    if (FAST && cbase + WCH <= n) {
      const StreamPos p0 = stream_locate(cbase, g);
      uint4 q[WCH_ROUNDS];
      uint2 kp[WCH_ROUNDS];
      float4 pa[WCH_ROUNDS], pb[WCH_ROUNDS];
      bool key[WCH_ROUNDS];
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) {
        const unsigned int off = it * 256 + lane * 8;
        const long long i0 = cbase + off;
        q[it] = *reinterpret_cast<const uint4 *>(body + i0);
        const StreamPos p = stream_advance(p0, off, FE);
        const int s = slot[p.f];
        key[it] = s < 0;
        if (s < 0) {
          kp[it] = *reinterpret_cast<const uint2 *>(key_plane + i0);
        } else {
          const float4 *pp = reinterpret_cast<const float4 *>(pool + (long long)s * g.pframe_elems + pool_ofs(p.r, g));
          pa[it] = pp[0];
          pb[it] = pp[1];
        }
      }
#pragma unroll
      for (int it = 0; it < WCH_ROUNDS; it++) {
        const long long i0 = cbase + it * 256 + lane * 8;
        int v[8];
        unpack8_i16(q[it], v);
        map8(sl, use_lut != 0, v);
        unsigned int loc[8], run = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          run += (unsigned int)v[k];
          loc[k] = run;
        }
        unsigned int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const unsigned int base = before + inc - run;
        before += __shfl_sync(0xffffffffu, inc, 31);
        int xs[8];
#pragma unroll
        for (int k = 0; k < 8; k++) xs[k] = (int)(int16_t)(xy - (base + loc[k]));
        if (x_out) *reinterpret_cast<uint4 *>(x_out + i0) = pack8_i16(xs);
        int P[8];
        if (key[it]) {
          unpack8_u8(kp[it], P);
        } else {
          P[0] = q255(pa[it].x); P[1] = q255(pa[it].y); P[2] = q255(pa[it].z); P[3] = q255(pa[it].w);
          P[4] = q255(pb[it].x); P[5] = q255(pb[it].y); P[6] = q255(pb[it].z); P[7] = q255(pb[it].w);
        }
        unsigned int o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = (unsigned int)min(max(P[k] - xs[k], 0), 255);   // :252-256,269
        uint2 w;
        w.x = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        w.y = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
        *reinterpret_cast<uint2 *>(out + i0) = w;
      }
    } else {
      // the last, partial chunk of a stream, and geometries whose rows are not multiples of 8 samples
#pragma unroll 1
      for (int it = 0; it < WCH_ROUNDS; it++) {
        const long long i0 = cbase + it * 256 + lane * 8;
        int v[8];
        load8_i16(body, i0, n, v);
        map8(sl, use_lut != 0, v);
        unsigned int loc[8], run = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          run += (i0 + k < n) ? (unsigned int)v[k] : 0u;
          loc[k] = run;
        }
        unsigned int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const unsigned int base = before + inc - run;
        before += __shfl_sync(0xffffffffu, inc, 31);
        int xs[8];
#pragma unroll
        for (int k = 0; k < 8; k++) xs[k] = (int)(int16_t)(xy - (base + loc[k]));
        if (x_out) store8_i16(x_out, i0, n, xs);
#pragma unroll 1
        for (int k = 0; k < 8; k++) {
          const long long i = i0 + k;
          if (i < n) {
            const long long f = i / g.frame_elems;
            const int s = slot[f];
            int P;
            if (s < 0) {
              P = key_plane[i];
            } else {
              const unsigned int r = (unsigned int)rem_in_frame(i, f, g.frame_elems);
              P = q255(pool[(long long)s * g.pframe_elems + pool_ofs(r, g)]);
            }
            out[i] = (uint8_t)min(max(P - xs[k], 0), 255);
          }
        }
      }
    }
  }
}

// persistent grid for the warp-chunk kernels: 8 warps per block, at most per_sm resident blocks per SM
template <typename K>
static int resident_blocks(K kernel) {   // 256-thread blocks of `kernel` that fit on one SM (registers / shared memory)
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 256, 0) != cudaSuccess || nb < 1) nb = 1;
  return nb;
}
static int chunk_grid(long long n, int per_sm) {
  const long long blocks = ((n + WCH - 1) / WCH + 7) / 8;
  const long long cap = (long long)tz::sm_count() * per_sm;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ------------------------------------------------------------------------------------------------ misc
// compress.py:138,176 / decompress.py:117,120: normalise through the 256-entry LUT and zero-pad.
__global__ void __launch_bounds__(256) pad_normalize_kernel(const uint8_t *__restrict__ frames,
                                                            const int32_t *__restrict__ frame_idx,
                                                            const float *__restrict__ lut, float *__restrict__ out,
                                                            int B, Geo g) {
  __shared__ float sl[256];
  sl[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  long long total = (long long)B * g.pframe_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long b = i / g.pframe_elems;
    int r = rem_in_frame(i, b, g.pframe_elems);
    int row = r / g.prow;
    int col = r - row * g.prow;
    float v = 0.0f;
    if (row < g.H && col < g.rowlen) {
      long long f = frame_idx ? (long long)frame_idx[b] : b;
      v = sl[frames[f * g.frame_elems + (long long)row * g.rowlen + col]];
    }
    out[i] = v;
  }
}

// compress.py:245-246: sum of squared error over the padded frame, float64, fixed reduction order.
__global__ void __launch_bounds__(1024) window_sse_kernel(const uint8_t *__restrict__ frames,
                                                          const int32_t *__restrict__ frame_idx,
                                                          const float *__restrict__ lut,
                                                          const float *__restrict__ pred, double *__restrict__ sse,
                                                          Geo g) {
  __shared__ float sl[256];
  __shared__ double red[1024];
  if (threadIdx.x < 256) sl[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const int b = blockIdx.x;
  const long long f = frame_idx ? (long long)frame_idx[b] : (long long)b;
  const float *p = pred + (long long)b * g.pframe_elems;
  double acc = 0.0;
  for (int i = threadIdx.x; i < (int)g.pframe_elems; i += 1024) {
    int row = i / g.prow;
    int col = i - row * g.prow;
    double a = 0.0;
    if (row < g.H && col < g.rowlen) a = (double)sl[frames[f * g.frame_elems + (long long)row * g.rowlen + col]];
    double d = __dsub_rn(a, (double)p[i]);
    acc = __dadd_rn(acc, __dmul_rn(d, d));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) sse[b] = red[0];
}

__global__ void __launch_bounds__(256) key_plane_kernel(const uint8_t *__restrict__ frames,
                                                        const uint8_t *__restrict__ is_key,
                                                        uint8_t *__restrict__ out, long long nt,
                                                        long long frame_bytes) {
  // grid.y = frame; 16-byte vector path when frame_bytes % 16 == 0
  long long f = blockIdx.y;
  const bool key = is_key[f] != 0;
  const uint8_t *src = frames + f * frame_bytes;
  uint8_t *dst = out + f * frame_bytes;
  if ((frame_bytes & 15) == 0) {
    long long nv = frame_bytes >> 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
         i += (long long)gridDim.x * blockDim.x) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (key) v = reinterpret_cast<const uint4 *>(src)[i];
      reinterpret_cast<uint4 *>(dst)[i] = v;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frame_bytes;
         i += (long long)gridDim.x * blockDim.x)
      dst[i] = key ? src[i] : (uint8_t)0;
  }
}

__global__ void __launch_bounds__(256) frames_nonzero_kernel(const uint8_t *__restrict__ plane,
                                                             uint8_t *__restrict__ nonzero,
                                                             long long frame_bytes) {
  long long f = blockIdx.x;
  const uint8_t *src = plane + f * frame_bytes;
  int any = 0;
  if ((frame_bytes & 15) == 0) {
    long long nv = frame_bytes >> 4;
    for (long long i = threadIdx.x; i < nv; i += blockDim.x) {
      uint4 v = reinterpret_cast<const uint4 *>(src)[i];
      any |= (v.x | v.y | v.z | v.w) != 0;
    }
  } else {
    for (long long i = threadIdx.x; i < frame_bytes; i += blockDim.x) any |= src[i] != 0;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) nonzero[f] = any ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ DWP on the device
// The dynamic-window scheduler (compress.py:214-266) for B chains that advance in lock step, without a host round
// trip per step: the inputs of the next PredNet step are gathered by chain state (key frame -> normalised + padded,
// compress.py:219; otherwise the chain's previous prediction, :222), and after the step one thread per chain takes the
// close decision of :245-263 and updates the state, the key flags and the frame -> pool-slot table.
template <typename PIX>
__global__ void __launch_bounds__(256) dwp_gather_kernel(const PIX *__restrict__ frames, const float *__restrict__ pool,
                                                         const int32_t *__restrict__ key,
                                                         const int32_t *__restrict__ idx,
                                                         const int32_t *__restrict__ last, float *__restrict__ X, Geo g,
                                                         float inv_unused) {
  const int b = blockIdx.y;
  const bool from_key = idx[b] == key[b] + 1;
  float *dst = X + (long long)b * g.pframe_elems;
  if (from_key) {
    const PIX *src = frames + (long long)key[b] * g.frame_elems;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)g.pframe_elems; i += gridDim.x * blockDim.x) {
      const int row = i / g.prow, col = i - row * g.prow;
      float v = 0.0f;
      if (row < g.H && col < g.rowlen) {
        const PIX s = src[(long long)row * g.rowlen + col];
        // compress.py:138: float32(sample) / pixel maximum (IEEE division == the 256-entry table of tz_pad_normalize)
        v = sizeof(PIX) == 1 ? __fdiv_rn((float)s, 255.0f) : __fdiv_rn((float)s, 65535.0f);
      }
      dst[i] = v;
    }
  } else {
    const float4 *src = reinterpret_cast<const float4 *>(pool + (long long)last[b] * g.pframe_elems);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)(g.pframe_elems >> 2); i += gridDim.x * blockDim.x)
      d4[i] = src[i];
  }
}

__global__ void dwp_update_kernel(const double *__restrict__ sse_step, int32_t *__restrict__ key,
                                  int32_t *__restrict__ idx, int32_t *__restrict__ last, double *__restrict__ sse,
                                  int32_t *__restrict__ cnt, int32_t *__restrict__ pred_slot,
                                  uint8_t *__restrict__ apply, uint8_t *__restrict__ is_key, int B, int slot0,
                                  double denom, int has_threshold, double threshold, int window, int p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double s = __dadd_rn(sse[b], sse_step[b]);
  const int c = cnt[b] + 1;
  const int i = idx[b];
  const double stop_point = __ddiv_rn(s, __dmul_rn((double)c, denom));                  // compress.py:245-246
  const bool closes = (has_threshold && stop_point > threshold) || (window > 0 && ((i - p) % window) == 0);   // :249
  if (closes) {            // :251-263: frame i becomes the next key, its prediction is dropped
    is_key[i] = 1;
    key[b] = i;
    sse[b] = 0.0;
    cnt[b] = 0;
    last[b] = -1;
    pred_slot[i] = -1;
    apply[i] = 0;
  } else {
    pred_slot[i] = slot0 + b;
    apply[i] = 1;
    last[b] = slot0 + b;
    sse[b] = s;
    cnt[b] = c;
  }
  idx[b] = i + 1;
}

// last residual of a shard: the one-element halo of the 1-D delta across shards (compress.py:75), left on the device
__global__ void last_residual_kernel(const uint8_t *__restrict__ frames, const float *__restrict__ pool,
                                     const int32_t *__restrict__ slot, Geo g, long long n, int32_t *__restrict__ out) {
  out[0] = resid_at(frames, pool, slot, g, n - 1);
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int tz_pad_normalize(const uint8_t *frames, const int32_t *frame_idx, const float *lut, float *out, int B,
                     int H, int W, int C, int Hp, int Wp, void *stream) {
  TZ_REQUIRE(frames && lut && out && B >= 0 && H > 0 && W > 0 && C > 0 && Hp >= H && Wp >= W,
             "tz_pad_normalize: bad arguments");
  if (B == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  int grid = stream_grid((long long)B * g.pframe_elems, 256, 8);
  pad_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, frame_idx, lut, out, B, g);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_residual(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, int16_t *x,
                long long nt, int H, int W, int C, int Hp, int Wp, void *stream) {
  TZ_REQUIRE(frames && pred_pool && pred_slot && x && nt >= 0 && H > 0 && W > 0 && C > 0 && Hp >= H && Wp >= W,
             "tz_residual: bad arguments");
  if (nt == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  long long n = nt * g.frame_elems;
  int grid = stream_grid((n + 7) / 8, 256, 8);
  if (fast_ok(g))
    residual_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, pred_pool, pred_slot, x, n, g);
  else
    residual_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, pred_pool, pred_slot, x, n, g);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_last_residual(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, long long nt, int H,
                     int W, int C, int Hp, int Wp, int32_t *out, void *stream) {
  TZ_REQUIRE(frames && pred_pool && pred_slot && out && nt > 0 && H > 0 && W > 0 && C > 0 && Hp >= H && Wp >= W,
             "tz_last_residual: bad arguments");
  Geo g = make_geo(H, W, C, Hp, Wp);
  last_residual_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(frames, pred_pool, pred_slot, g, nt * g.frame_elems, out);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_error_bound(const uint8_t *frames, int16_t *x, const uint8_t *apply, long long nt, int H, int W,
                   int C, int mode, double b0, double b1, void *stream) {
  TZ_REQUIRE(frames && x && apply && nt >= 0 && H > 0 && W > 0 && C > 0, "tz_error_bound: bad arguments");
  TZ_REQUIRE(mode >= TZ_MODE_ABS && mode <= TZ_MODE_PWREL, "tz_error_bound: unknown mode %d", mode);
  if (b0 == 0.0) return TZ_OK;                                   // compress.py:24
  if (mode == TZ_MODE_ABSREL && b1 == 0.0) return TZ_OK;         // compress.py:35
  if (nt == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, H, W);
  long long planes = nt * C;
  TZ_REQUIRE(planes < 2147483647LL, "tz_error_bound: too many planes (%lld)", planes);
  static const bool legacy = getenv("TZ_EB_LEGACY") != nullptr;   // A/B switch: the warp-serial kernel for every mode
  if (mode != TZ_MODE_PWREL && !legacy) {
    // Tile / CTA shape measured on B200 (2700 planes of 20480 samples, abs 2): <1024, 128> 0.63 ms, <512, 128> 0.67,
    // <1024, 256> 0.69, <512, 64> 0.81, <1024, 64> 0.88.
    error_bound_tiles_kernel<1024, 128><<<(unsigned)planes, 128, 0, (cudaStream_t)stream>>>(frames, x, apply, nt, g, mode,
                                                                                           b0, b1);
  } else {
    const int threads = 128;   // 4 warps = 4 planes per block
    long long blocks = (planes * 32 + threads - 1) / threads;
    error_bound_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(frames, x, apply, nt, g, mode, b0, b1);
  }
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_delta_hist(const int16_t *x, long long n, int has_prev, const int32_t *prev_x, unsigned long long *hist,
                  unsigned long long *overflow, void *stream) {
  TZ_REQUIRE(x && hist && overflow && n >= 0, "tz_delta_hist: bad arguments");
  TZ_REQUIRE(has_prev >= 0 && has_prev <= 2 && (has_prev != 1 || prev_x), "tz_delta_hist: bad has_prev / prev_x");
  if (n == 0) return TZ_OK;
  Geo g = make_geo(1, 1, 1, 1, 1);
  static const int per_sm = resident_blocks(delta_hist_kernel<0>);
  int grid = chunk_grid(n, per_sm);
  delta_hist_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, nullptr, g, n, has_prev, prev_x,
                                                              hist, overflow);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_build_table(const unsigned long long *hist, int16_t *table, int16_t *lut, int32_t *meta, void *stream) {
  TZ_REQUIRE(hist && table && lut && meta, "tz_build_table: null argument");
  build_table_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(hist, table, lut, meta);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_delta_rank(const int16_t *x, long long n, int has_prev, const int32_t *prev_x, const int16_t *lut,
                  int16_t *out, void *stream) {
  TZ_REQUIRE(x && out && n >= 0, "tz_delta_rank: bad arguments");
  TZ_REQUIRE(has_prev >= 0 && has_prev <= 2 && (has_prev != 1 || prev_x), "tz_delta_rank: bad has_prev / prev_x");
  if (n == 0) return TZ_OK;
  Geo g = make_geo(1, 1, 1, 1, 1);
  static const int per_sm = resident_blocks(delta_rank_kernel<0>);
  int grid = chunk_grid(n, per_sm);
  delta_rank_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, nullptr, g, n, has_prev, prev_x,
                                                              lut, out);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_encode_lossless(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, long long nt,
                       int H, int W, int C, int Hp, int Wp, int has_prev, const int32_t *prev_x, int pass,
                       unsigned long long *hist, unsigned long long *overflow, const int16_t *lut,
                       int16_t *out, void *stream) {
  TZ_REQUIRE(frames && pred_pool && pred_slot && nt >= 0 && H > 0 && W > 0 && C > 0 && Hp >= H && Wp >= W,
             "tz_encode_lossless: bad arguments");
  TZ_REQUIRE(has_prev >= 0 && has_prev <= 1 && (has_prev != 1 || prev_x), "tz_encode_lossless: bad has_prev / prev_x");
  TZ_REQUIRE(pass == 0 || pass == 1, "tz_encode_lossless: pass must be 0 or 1");
  if (nt == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  long long n = nt * g.frame_elems;
  int grid = chunk_grid(n, 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (pass == 0) {
    TZ_REQUIRE(hist && overflow, "tz_encode_lossless: pass 0 needs hist and overflow");
    if (fast_ok(g))
      delta_hist_kernel<2><<<grid, 256, 0, st>>>(nullptr, frames, pred_pool, pred_slot, g, n, has_prev, prev_x, hist, overflow);
    else
      delta_hist_kernel<1><<<grid, 256, 0, st>>>(nullptr, frames, pred_pool, pred_slot, g, n, has_prev, prev_x, hist, overflow);
  } else {
    TZ_REQUIRE(out, "tz_encode_lossless: pass 1 needs out");
    if (fast_ok(g))
      delta_rank_kernel<2><<<grid, 256, 0, st>>>(nullptr, frames, pred_pool, pred_slot, g, n, has_prev, prev_x, lut, out);
    else
      delta_rank_kernel<1><<<grid, 256, 0, st>>>(nullptr, frames, pred_pool, pred_slot, g, n, has_prev, prev_x, lut, out);
  }
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

long long tz_reconstruct_workspace_bytes(long long n) {
  long long nchunks = (n + WCH - 1) / WCH;
  return ((nchunks + 7) / 8 * 8 + 8) * (long long)sizeof(unsigned int);
}

int tz_reconstruct(const int16_t *body, long long nt, int H, int W, int C, int Hp, int Wp, int table_len,
                   const int16_t *rank_lut, int first_mode, int first_x, const float *pred_pool,
                   const int32_t *pred_slot, const uint8_t *key_plane, uint8_t *out, int16_t *x_out,
                   void *workspace, void *stream) {
  TZ_REQUIRE(body && pred_pool && pred_slot && key_plane && out && workspace && nt >= 0 && H > 0 && W > 0 &&
                 C > 0 && Hp >= H && Wp >= W,
             "tz_reconstruct: bad arguments");
  TZ_REQUIRE(table_len < 0 || rank_lut, "tz_reconstruct: rank_lut required when table_len >= 0");
  if (nt == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  TZ_REQUIRE(g.frame_elems < (1LL << 31), "tz_reconstruct: frame too large");
  long long n = nt * g.frame_elems;
  long long nchunks = (n + WCH - 1) / WCH;
  long long npad = (nchunks + 7) / 8 * 8;
  unsigned int *sums = (unsigned int *)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  int use_lut = table_len >= 0;
  static const int sum_per_sm = resident_blocks(decode_wsum_kernel);
  decode_wsum_kernel<<<chunk_grid(n, sum_per_sm), 256, 0, st>>>(body, n, npad, use_lut, rank_lut, sums);
  TZ_CHECK_LAUNCH();
  decode_scan8_kernel<<<1, 1024, 0, st>>>(sums, npad);
  TZ_CHECK_LAUNCH();
  if (fast_ok(g))
    decode_wrecon_kernel<true><<<chunk_grid(n, 2), 256, 0, st>>>(
        body, n, nchunks, use_lut, rank_lut, sums, first_mode, first_x, pred_pool, pred_slot, key_plane, out, x_out, g);
  else
    decode_wrecon_kernel<false><<<chunk_grid(n, 2), 256, 0, st>>>(
        body, n, nchunks, use_lut, rank_lut, sums, first_mode, first_x, pred_pool, pred_slot, key_plane, out, x_out, g);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_window_sse(const uint8_t *frames, const int32_t *frame_idx, const float *lut, const float *pred,
                  double *sse, int B, int H, int W, int C, int Hp, int Wp, void *stream) {
  TZ_REQUIRE(frames && lut && pred && sse && B >= 0 && H > 0 && W > 0 && C > 0 && Hp >= H && Wp >= W,
             "tz_window_sse: bad arguments");
  if (B == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  TZ_REQUIRE(g.pframe_elems < 2147483647LL, "tz_window_sse: frame too large");
  window_sse_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(frames, frame_idx, lut, pred, sse, g);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_dwp_gather(const void *frames, int bits, const float *pred_pool, const int32_t *key, const int32_t *idx,
                  const int32_t *last, float *X, int B, int H, int W, int C, int Hp, int Wp, void *stream) {
  TZ_REQUIRE(frames && pred_pool && key && idx && last && X && B >= 0 && H > 0 && W > 0 && C > 0 && Hp >= H && Wp >= W &&
                 (bits == 8 || bits == 16), "tz_dwp_gather: bad arguments");
  if (B == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  TZ_REQUIRE(g.pframe_elems < 2147483647LL && (g.pframe_elems % 4) == 0 && B <= 65535, "tz_dwp_gather: frame too large");
  int gx = (int)((g.pframe_elems / 4 + 255) / 256);
  if (gx > 64) gx = 64;
  dim3 grid(gx, B);
  if (bits == 8)
    dwp_gather_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)frames, pred_pool, key, idx, last, X, g, 0.f);
  else
    dwp_gather_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t *)frames, pred_pool, key, idx, last, X, g, 0.f);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_dwp_update(const double *sse_step, int32_t *key, int32_t *idx, int32_t *last, double *sse, int32_t *cnt,
                  int32_t *pred_slot, uint8_t *apply, uint8_t *is_key, int B, int slot0, double denom,
                  int has_threshold, double threshold, int window, int p, void *stream) {
  TZ_REQUIRE(sse_step && key && idx && last && sse && cnt && pred_slot && apply && is_key && B >= 0 && denom > 0.0,
             "tz_dwp_update: bad arguments");
  if (B == 0) return TZ_OK;
  dwp_update_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sse_step, key, idx, last, sse, cnt, pred_slot, apply,
                                                                      is_key, B, slot0, denom, has_threshold, threshold,
                                                                      window, p);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_key_plane(const uint8_t *frames, const uint8_t *is_key, uint8_t *out, long long nt,
                 long long frame_bytes, void *stream) {
  TZ_REQUIRE(frames && is_key && out && nt >= 0 && frame_bytes > 0 && nt < 65536LL * 32768LL,
             "tz_key_plane: bad arguments");
  if (nt == 0) return TZ_OK;
  // grid.y is limited to 65535: loop over slabs of frames
  for (long long f0 = 0; f0 < nt; f0 += 65535) {
    long long nf = nt - f0 < 65535 ? nt - f0 : 65535;
    int gx = (int)((frame_bytes / 16 + 255) / 256);
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    dim3 grid(gx, (unsigned)nf);
    key_plane_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames + f0 * frame_bytes, is_key + f0,
                                                            out + f0 * frame_bytes, nf, frame_bytes);
    TZ_CHECK_LAUNCH();
  }
  return TZ_OK;
}

int tz_frames_nonzero(const uint8_t *key_plane, uint8_t *nonzero, long long nt, long long frame_bytes,
                      void *stream) {
  TZ_REQUIRE(key_plane && nonzero && nt >= 0 && frame_bytes > 0 && nt < 2147483647LL,
             "tz_frames_nonzero: bad arguments");
  if (nt == 0) return TZ_OK;
  frames_nonzero_kernel<<<(unsigned)nt, 256, 0, (cudaStream_t)stream>>>(key_plane, nonzero, frame_bytes);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_memcpy2d_async(void *dst, long long dpitch, const void *src, long long spitch, long long width,
                      long long height, void *stream) {
  TZ_REQUIRE(dst && src && width >= 0 && height >= 0 && dpitch >= width && spitch >= width,
             "tz_memcpy2d_async: bad arguments");
  if (width == 0 || height == 0) return TZ_OK;
  TZ_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height,
                                  cudaMemcpyDefault, (cudaStream_t)stream));
  return TZ_OK;
}

}  // extern "C"
