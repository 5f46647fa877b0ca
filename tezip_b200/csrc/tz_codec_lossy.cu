// tezip_b200 -- the lossy encode as ONE data pass: residual (compress.py:293-314) + error-bound quantisation
// (compress.py:23-70, :315-319) + 1-D delta and symbol histogram (compress.py:73-77, :348-355) for 8-bit samples.
// Reads frames (1 B) and predictions (4 B) once, writes the quantised residual x (2 B) once; the rank map
// (tz_delta_rank) reads x once more.  Replaces the four-kernel chain residual -> error_bound -> delta_hist (17 B of
// HBM traffic per sample, and an error_bound kernel that was bound by instruction issue: ~250 warp instructions per 32
// samples).
//
// error_bound is a greedy serial scan: a segment is flushed when the next element would empty the running interval
// intersection -- for a plane-wide bound E and integer residuals: when max - min of the segment would exceed
// G = floor(2E) (the exact integer form of the reference's float64 test; conditions and proof in tz_codec.cu).
// Parallel form used here, one warp per (frame, channel) plane, tiles of 32 x L elements:
//   speculate  every lane scans its L consecutive elements from a FRESH state and records where it breaks;
//   correct    a greedy scan started anywhere re-synchronises with the true one as soon as both break at the same
//              element.  So lane k re-walks its range from the state that really enters it (its left neighbour's end
//              state) only until it hits one of its own recorded breaks; a range that the entering segment swallows
//              whole is absorbed in O(1) through its (min, max) summary.  All lanes do this at once from their
//              neighbours' current end states, and the round is repeated while some end state still changes: one
//              round for noisy residuals (measured: ~18 steps), at most 31 (then it is the serial scan);
//   values     forward pass parks each closing segment's value trunc((u+l)/2) on its last element, backward pass
//              spreads it; a segment still open at the end of a tile is written when it closes.
// scripts/eb_spec_emulate.py is a lane-level numpy emulation of exactly this, checked against the oracle.
// Planes whose bound defeats the integer shortcut take the IEEE-double warp scan (eb_plane_warp<EbDbl>), pwrel and
// odd geometries keep the separate kernels of tz_codec.cu (the host side chooses).
#include "tz_codec.cuh"

namespace {

constexpr int LZ_L = 32;                   // elements per lane and tile
constexpr int LZ_T = 32 * LZ_L;            // elements per tile
constexpr int LZ_STRIDE = 34;              // halfwords per row j of the tile buffer: 17 words, odd -> rows hit distinct banks
constexpr int LZ_MAXC = 4;                 // channels (one warp each)
constexpr int LZ_BIG = 1 << 30;
constexpr int LZ_WIN_LO = 1088, LZ_WIN = 1024;   // shared histogram window: symbols 1088..2111 (valid streams: 1090..2110)

struct LzSmem {
  int16_t e[LZ_MAXC][LZ_L * LZ_STRIDE];    // e[c][j * STRIDE + k]: element j of lane k's range
  unsigned int hist[LZ_WIN];
  int red[2 * LZ_MAXC * LZ_MAXC];
  int fallback;
  unsigned int ticket;
};

template <bool EXACT>
__device__ __forceinline__ int lz_mid(int mn, int mx, double E) {
  if (EXACT) return (mn + mx) / 2;         // exact sums: trunc(((a+E)+(b-E))/2) == truncating integer division
  return (int)(long long)__dmul_rn(__dadd_rn(__dadd_rn((double)mn, E), __dsub_rn((double)mx, E)), 0.5);   // compress.py:61
}

// One tile of one plane, executed by one warp.  e: this warp's tile buffer; cnt: valid elements of this lane's range
// (LZ_L except in the last tile of a plane); carry (mn, mx) / head: state and first plane index of the segment that
// is open on entry (warp-uniform); xplane: global x of this plane (element i at xplane[i * C]); t0: plane index of
// the tile's first element.  The lane's LZ_L elements live in REGISTERS for all four passes (every loop below has a
// compile-time trip count and is fully unrolled; invalid tail elements are predicated off), so the passes cost ~8
// integer instructions per element and no shared-memory traffic beyond one load and one store per element.
template <bool EXACT, bool FT /* full tile: every lane owns LZ_L valid elements (no tail predicates) */>
__device__ __forceinline__ void eb_spec_tile(int16_t *__restrict__ e, int cnt, int lane, int &carry_mn, int &carry_mx,
                                             int &head, int16_t *__restrict__ xplane, int C, int t0, int n_tile,
                                             int G, double E) {
  constexpr unsigned FULL = 0xffffffffu;
  int v[LZ_L];
  // ---- speculate
  int mn = lane == 0 ? carry_mn : LZ_BIG, mx = lane == 0 ? carry_mx : -LZ_BIG;
  int smn = LZ_BIG, smx = -LZ_BIG;
  unsigned fl = 0;
  const unsigned valid = (FT || cnt >= 32) ? FULL : ((1u << cnt) - 1u);   // bit j: element j of this lane's range exists
#pragma unroll
  for (int j = 0; j < LZ_L; j++) {
    v[j] = e[j * LZ_STRIDE + lane];
    const bool in = FT || ((valid >> j) & 1u);
    const int d = v[j];
    smn = in ? min(smn, d) : smn;
    smx = in ? max(smx, d) : smx;
    const int nmn = min(mn, d), nmx = max(mx, d);
    const bool br = in && (nmx - nmn > G);
    fl |= (unsigned)br << j;
    mn = br ? d : (in ? nmn : mn);
    mx = br ? d : (in ? nmx : mx);
  }
  int end_mn = mn, end_mx = mx;
  int in_mn = carry_mn, in_mx = carry_mx;   // lane 0: the true entering state; others: set in round 1
  // ---- correct
  bool active = lane > 0;
  while (__any_sync(FULL, active)) {
    const int pmn = __shfl_up_sync(FULL, end_mn, 1), pmx = __shfl_up_sync(FULL, end_mx, 1);
    bool changed = false;
    const int jmn = min(pmn, smn), jmx = max(pmx, smx);
    const bool swallow = jmx - jmn <= G;      // the entering segment swallows the whole range
    // predicated walk (all lanes step together; a lane stops updating once it is synchronised or inactive)
    int a = pmn, b = pmx;
    unsigned nf = 0, keep = 0;
    bool walking = active && !swallow;
    if (__any_sync(FULL, walking)) {
#pragma unroll
      for (int j = 0; j < LZ_L; j++) {
        const bool in = walking && (FT || ((valid >> j) & 1u));
        const int d = v[j];
        const int na = min(a, d), nb = max(b, d);
        const bool br = in && (nb - na > G);
        const bool sync_here = br && ((fl >> j) & 1u);   // both scans break here: from here on they are the same scan
        keep = sync_here ? (fl & ~((1u << j) - 1u)) : keep;
        walking = walking && !sync_here;
        const bool take = br && !sync_here;
        nf |= (unsigned)take << j;
        a = take ? d : (in && !br ? na : a);
        b = take ? d : (in && !br ? nb : b);
      }
    }
    if (active) {
      in_mn = pmn;
      in_mx = pmx;
      int nmn_end, nmx_end;
      if (swallow) {
        fl = 0;
        nmn_end = jmn;
        nmx_end = jmx;
      } else if (keep) {            // synchronised (keep holds at least the bit of the common break)
        fl = nf | keep;
        nmn_end = end_mn;
        nmx_end = end_mx;
      } else {
        fl = nf;
        nmn_end = a;
        nmx_end = b;
      }
      changed = (nmn_end != end_mn) | (nmx_end != end_mx);
      end_mn = nmn_end;
      end_mx = nmx_end;
    }
    active = (__shfl_up_sync(FULL, (int)changed, 1) != 0) && lane > 0;
  }
  // ---- values: forward (park each closing value on the segment's last element)
  int fc = 0;
  bool fc_valid = false;
  mn = in_mn;
  mx = in_mx;
#pragma unroll
  for (int j = 0; j < LZ_L; j++) {
    const int d = v[j];
    const bool br = (fl >> j) & 1u;            // (only valid elements carry a flag)
    const bool in = FT || ((valid >> j) & 1u);
    const int q = lz_mid<EXACT>(mn, mx, E);
    fc = (br && !fc_valid) ? q : fc;
    fc_valid = fc_valid || br;
    if (j >= 1) v[j - 1] = br ? q : v[j - 1];
    mn = br ? d : (in ? min(mn, d) : mn);
    mx = br ? d : (in ? max(mx, d) : mx);
  }
  const unsigned mask = __ballot_sync(FULL, fc_valid);
  const unsigned later = (lane == 31) ? 0u : (mask & ~((2u << lane) - 1u));   // lanes after this one that close a segment
  const int src = later ? __ffs(later) - 1 : 0;
  const int tail = __shfl_sync(FULL, fc, src);
  bool curv = later != 0;
  if (mask) {   // the segment that was open on entry closes in this tile: write its part that lies in earlier tiles
    const int q0 = __shfl_sync(FULL, fc, __ffs(mask) - 1);
    for (int i = head + lane; i < t0; i += 32) xplane[(long long)i * C] = (int16_t)q0;
  }
  // ---- values: backward (spread); elements of a segment that stays open keep their residual and are rewritten later
  int cur = tail;
#pragma unroll
  for (int j = LZ_L - 1; j >= 0; j--) {
    const bool nb = (j + 1 < LZ_L) && ((fl >> (j + 1)) & 1u);   // the next element starts a segment: v[j] holds this one's value
    cur = nb ? v[j] : cur;
    curv = curv || nb;
    v[j] = curv ? cur : v[j];
    e[j * LZ_STRIDE + lane] = (int16_t)v[j];
  }
  // ---- carry
  const unsigned has = __ballot_sync(FULL, fl != 0);
  if (has) {
    const int kmax = 31 - __clz((int)has);
    const unsigned flk = __shfl_sync(FULL, fl, kmax);
    head = t0 + kmax * LZ_L + (31 - __clz((int)flk));
  }
  const int klast = (n_tile - 1) / LZ_L;
  carry_mn = __shfl_sync(FULL, end_mn, klast);
  carry_mx = __shfl_sync(FULL, end_mx, klast);
}

__device__ __forceinline__ void lz_hist_add(LzSmem &sm, unsigned long long *__restrict__ hist, unsigned int &ovf, int s,
                                            unsigned int cnt) {
  const unsigned int w = (unsigned int)(s - LZ_WIN_LO);
  if (w < (unsigned int)LZ_WIN) atomicAdd(&sm.hist[w], cnt);
  else if ((unsigned int)s < (unsigned int)TZ_HIST_BINS) atomicAdd(&hist[s], (unsigned long long)cnt);
  else ovf += cnt;
}

// Residuals of the 8 consecutive samples at in-frame stream offset i (a multiple of 8; rowlen % 8 == 0 keeps them in
// one row).  32-bit index arithmetic only: fp / pp are the frame's and its prediction's base pointers.
struct LzGroup {
  uint2 a;
  float4 p0, p1;
};
__device__ __forceinline__ void lz_load(LzGroup &q, const uint8_t *__restrict__ fp, const float *__restrict__ pp,
                                        const Geo &g, int i) {
  int pofs = i;
  if (g.prow != g.rowlen) {   // padded predictions: (row, col) of the cropped frame -> padded offset
    const int row = i / g.rowlen;
    pofs = row * g.prow + (i - row * g.rowlen);
  }
  q.a = *reinterpret_cast<const uint2 *>(fp + i);
  const float4 *src = reinterpret_cast<const float4 *>(pp + pofs);
  q.p0 = src[0];
  q.p1 = src[1];
}
__device__ __forceinline__ void lz_resid(const LzGroup &q, int v[8]) {   // compress.py:307,310-313
  v[0] = q255(q.p0.x) - (int)(q.a.x & 0xff);
  v[1] = q255(q.p0.y) - (int)((q.a.x >> 8) & 0xff);
  v[2] = q255(q.p0.z) - (int)((q.a.x >> 16) & 0xff);
  v[3] = q255(q.p0.w) - (int)(q.a.x >> 24);
  v[4] = q255(q.p1.x) - (int)(q.a.y & 0xff);
  v[5] = q255(q.p1.y) - (int)((q.a.y >> 8) & 0xff);
  v[6] = q255(q.p1.z) - (int)((q.a.y >> 16) & 0xff);
  v[7] = q255(q.p1.w) - (int)(q.a.y >> 24);
}
__device__ __forceinline__ uint4 lz_pack(const int v[8]) {
  uint4 a;
  a.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
  a.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
  a.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
  a.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
  return a;
}

// blockDim = 32 * C (C = channels, a template parameter so that the pixel / channel split of a stream offset is a
// multiply-shift).  Frames are dealt to CTAs round-robin (persistent grid).
// has_prev: 0 first shard (y[0] = x[0]), 1 *prev_x, 3 "the first element of the stream is accounted for by the
// caller" (sharded runs: the halo arrives after this kernel has been queued).
template <int C>
__global__ void __launch_bounds__(32 * C, 21 / C) lossy_fused_kernel(
    const uint8_t *__restrict__ frames, const float *__restrict__ pool, const int32_t *__restrict__ slot,
    const uint8_t *__restrict__ apply, int16_t *__restrict__ x, long long nt, Geo g, int mode, double b0, double b1,
    int has_prev, const int32_t *__restrict__ prev_x, unsigned long long *__restrict__ hist,
    unsigned long long *__restrict__ overflow, unsigned int *__restrict__ counter) {
  __shared__ LzSmem sm;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NTHR = 32 * C;
  constexpr int GPT = LZ_T * C / (8 * NTHR);   // 8-sample groups per thread and full tile (= 4)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = g.H * g.W;                       // elements per plane
  const int FE = (int)g.frame_elems;
  for (int i = tid; i < LZ_WIN; i += NTHR) sm.hist[i] = 0;
  unsigned int ovf = 0;
  __syncthreads();
  for (long long f = blockIdx.x; f < nt; f += gridDim.x) {
    int16_t *xf = x + f * FE;
    const int s = slot[f];
    if (s < 0) {
      // first frame of a window: x = 0 (compress.py:314); every delta inside the frame is 0
      for (int i = tid * 8; i < FE; i += NTHR * 8) *reinterpret_cast<uint4 *>(xf + i) = make_uint4(0, 0, 0, 0);
      if (tid == 0) atomicAdd(&sm.hist[TZ_SYMBOL_OFFSET - LZ_WIN_LO], (unsigned int)(FE - 1));
      continue;
    }
    const uint8_t *fp = frames + f * FE;
    const float *pp = pool + (long long)s * g.pframe_elems;
    const bool eb_on = apply[f] != 0;
    double E = fabs(b0);                                                   // compress.py:29
    if (eb_on && mode != TZ_MODE_ABS) {
      // rel / absrel: max - min of the ORIGINAL plane (compress.py:31-33,36-43), one cooperative pass over the frame
      int mx[C], mn[C];
#pragma unroll
      for (int c = 0; c < C; c++) { mx[c] = 0; mn[c] = 255; }
      for (int p = tid; p < n; p += NTHR)
#pragma unroll
        for (int c = 0; c < C; c++) {
          const int v = fp[p * C + c];
          mx[c] = max(mx[c], v);
          mn[c] = min(mn[c], v);
        }
#pragma unroll
      for (int c = 0; c < C; c++) {
        mx[c] = __reduce_max_sync(FULL, mx[c]);
        mn[c] = __reduce_min_sync(FULL, mn[c]);
        if (lane == 0) {
          sm.red[(2 * c) * LZ_MAXC + warp] = mx[c];
          sm.red[(2 * c + 1) * LZ_MAXC + warp] = mn[c];
        }
      }
      __syncthreads();
      int pmx = 0, pmn = 255;   // this warp's plane = channel `warp`
      for (int w = 0; w < C; w++) {
        pmx = max(pmx, sm.red[(2 * warp) * LZ_MAXC + w]);
        pmn = min(pmn, sm.red[(2 * warp + 1) * LZ_MAXC + w]);
      }
      if (mode == TZ_MODE_REL) {
        E = __dmul_rn((double)(pmx - pmn), b0);                            // :33
      } else {
        const double a = fabs(b0), r = __dmul_rn((double)(pmx - pmn), b1); // :38-39
        E = (a < r) ? a : r;                                               // :40-43
      }
    }
    const double twoE = E + E;
    const double sc = E * 68719476736.0;   // 2^36
    const bool exact = (E < 4096.0) && (sc == floor(sc));
    const bool clear = fabs(twoE - rint(twoE)) > 1e-6;
    const bool shortcut = E >= 0.0 && (exact || clear);
    const int G = twoE >= 70000.0 ? 70000 : (int)floor(twoE);
    if (tid == 0) sm.fallback = 0;
    __syncthreads();
    if (eb_on && !shortcut && lane == 0) sm.fallback = 1;   // (benign race: every writer stores 1)
    __syncthreads();
    const bool spec = eb_on && !sm.fallback;
    if (!spec) {
      // residual only (warm-up frames, compress.py:315), or the IEEE-double scan for planes whose bound defeats the
      // integer shortcut: residual to global, then one warp per plane on the strided plane
      for (int i = tid * 8; i < FE; i += NTHR * 8) {
        LzGroup q;
        int v[8];
        lz_load(q, fp, pp, g, i);
        lz_resid(q, v);
        *reinterpret_cast<uint4 *>(xf + i) = lz_pack(v);
      }
      __syncthreads();
      if (eb_on) eb_plane_warp<EbDbl>(fp + warp, xf + warp, n, C, false, b0, E, 0);
    } else {
      int carry_mn = LZ_BIG, carry_mx = -LZ_BIG, head = 0;
      int16_t *e = sm.e[warp];
      int16_t *xplane = xf + warp;
      for (int t0 = 0; t0 < n; t0 += LZ_T) {
        const int n_tile = min(LZ_T, n - t0);
        const int s0 = t0 * C, s1 = (t0 + n_tile) * C;   // stream range inside the frame (multiples of 8)
        // A: residuals of pixels [t0, t0 + n_tile) x C channels -> per-channel tile buffers.  All loads of the tile
        // are issued before the first use (GPT independent 8-sample groups per thread).
        {
          LzGroup q[GPT];
#pragma unroll
          for (int u = 0; u < GPT; u++) {
            const int i = s0 + (u * NTHR + tid) * 8;
            if (i < s1) lz_load(q[u], fp, pp, g, i);
          }
#pragma unroll
          for (int u = 0; u < GPT; u++) {
            const int i = s0 + (u * NTHR + tid) * 8;
            if (i < s1) {
              int v[8];
              lz_resid(q[u], v);
              int p = (i - s0) / C, c = (i - s0) - p * C;
#pragma unroll
              for (int m = 0; m < 8; m++) {
                sm.e[c][(p % LZ_L) * LZ_STRIDE + (p / LZ_L)] = (int16_t)v[m];
                if (++c == C) { c = 0; p++; }
              }
            }
          }
        }
        __syncthreads();
        // B: one warp per channel plane
        {
          const int cnt = max(0, min(LZ_L, n_tile - lane * LZ_L));
          if (n_tile == LZ_T) {
            if (exact) eb_spec_tile<true, true>(e, cnt, lane, carry_mn, carry_mx, head, xplane, C, t0, n_tile, G, E);
            else eb_spec_tile<false, true>(e, cnt, lane, carry_mn, carry_mx, head, xplane, C, t0, n_tile, G, E);
          } else {
            if (exact) eb_spec_tile<true, false>(e, cnt, lane, carry_mn, carry_mx, head, xplane, C, t0, n_tile, G, E);
            else eb_spec_tile<false, false>(e, cnt, lane, carry_mn, carry_mx, head, xplane, C, t0, n_tile, G, E);
          }
        }
        __syncthreads();
        // C: tile buffers -> global x (16-byte stores)
#pragma unroll
        for (int u = 0; u < GPT; u++) {
          const int i = s0 + (u * NTHR + tid) * 8;
          if (i < s1) {
            int v[8];
            int p = (i - s0) / C, c = (i - s0) - p * C;
#pragma unroll
            for (int m = 0; m < 8; m++) {
              v[m] = (int)sm.e[c][(p % LZ_L) * LZ_STRIDE + (p / LZ_L)];
              if (++c == C) { c = 0; p++; }
            }
            *reinterpret_cast<uint4 *>(xf + i) = lz_pack(v);
          }
        }
        __syncthreads();
      }
      {   // compress.py:67: the segment still open at the end of the plane
        const int q = exact ? lz_mid<true>(carry_mn, carry_mx, E) : lz_mid<false>(carry_mn, carry_mx, E);
        for (int i = head + lane; i < n; i += 32) xplane[(long long)i * C] = (int16_t)q;
      }
    }
    __syncthreads();
    // D: delta (compress.py:75) + histogram of this frame from its final x (just written by this CTA: L2 hits).  The
    // frame's first element needs the last x of the previous frame (another CTA): the last CTA to finish adds those.
    for (int i0 = tid * 8; i0 < FE; i0 += NTHR * 8 * 4) {   // four independent 8-sample groups per thread in flight
      uint4 a4[4];
      int pv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * NTHR * 8;
        if (i < FE) {
          a4[u] = __ldcg(reinterpret_cast<const uint4 *>(xf + i));
          pv[u] = (i > 0) ? (int)__ldcg(xf + i - 1) : 0;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * NTHR * 8;
        if (i < FE) {
          const uint4 a = a4[u];
          int v[8];
          v[0] = (int16_t)(a.x & 0xffff); v[1] = (int16_t)(a.x >> 16);
          v[2] = (int16_t)(a.y & 0xffff); v[3] = (int16_t)(a.y >> 16);
          v[4] = (int16_t)(a.z & 0xffff); v[5] = (int16_t)(a.z >> 16);
          v[6] = (int16_t)(a.w & 0xffff); v[7] = (int16_t)(a.w >> 16);
          int prev = pv[u];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int y = (int)(int16_t)(prev - v[k]);
            const int sy = (int)(int16_t)(TZ_SYMBOL_OFFSET - y);              // :348 int16 arithmetic
            prev = v[k];
            if (i + k > 0) {
              const unsigned int w = (unsigned int)(sy - LZ_WIN_LO);
              if (w < (unsigned int)LZ_WIN) atomicAdd(&sm.hist[w], 1u);
              else lz_hist_add(sm, hist, ovf, sy, 1u);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = tid; i < LZ_WIN; i += NTHR) {
    const unsigned int c = sm.hist[i];
    if (c) atomicAdd(&hist[LZ_WIN_LO + i], (unsigned long long)c);
  }
  if (ovf) atomicAdd(overflow, (unsigned long long)ovf);
  // ---- the last CTA to finish adds the first symbol of every frame
  __threadfence();
  __syncthreads();
  if (tid == 0) sm.ticket = atomicAdd(counter, 1u);
  __syncthreads();
  if (sm.ticket != gridDim.x - 1) return;
  __threadfence();
  for (long long f = tid; f < nt; f += NTHR) {
    const int cur = (int)__ldcg(x + f * FE);
    int y;
    if (f == 0) {
      if (has_prev == 3) continue;
      y = (has_prev == 1) ? (int)(int16_t)(*prev_x - cur) : cur;
    } else {
      y = (int)(int16_t)((int)__ldcg(x + f * FE - 1) - cur);
    }
    const int sy = (int)(int16_t)(TZ_SYMBOL_OFFSET - y);
    if ((unsigned int)sy < (unsigned int)TZ_HIST_BINS) atomicAdd(&hist[sy], 1ull);
    else atomicAdd(overflow, 1ull);
  }
}

}  // namespace

extern "C" {

int tz_encode_lossy_supported(int H, int W, int C, int mode) {
  return (mode == TZ_MODE_ABS || mode == TZ_MODE_REL || mode == TZ_MODE_ABSREL) && C >= 1 && C <= LZ_MAXC &&
         ((W * C) % 8) == 0 && (long long)H * W * C < (1LL << 31);
}

int tz_encode_lossy(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, const uint8_t *apply,
                    int16_t *x, long long nt, int H, int W, int C, int Hp, int Wp, int mode, double b0, double b1,
                    int has_prev, const int32_t *prev_x, unsigned long long *hist, unsigned long long *overflow,
                    unsigned int *counter, void *stream) {
  TZ_REQUIRE(frames && pred_pool && pred_slot && apply && x && hist && overflow && counter && nt >= 0 && H > 0 && W > 0 &&
                 Hp >= H && Wp >= W,
             "tz_encode_lossy: bad arguments");
  TZ_REQUIRE(tz_encode_lossy_supported(H, W, C, mode), "tz_encode_lossy: unsupported geometry or mode (see "
             "tz_encode_lossy_supported); use tz_residual + tz_error_bound + tz_delta_hist");
  TZ_REQUIRE(b0 != 0.0 && !(mode == TZ_MODE_ABSREL && b1 == 0.0), "tz_encode_lossy: the bound is lossless; use tz_encode_lossless");
  TZ_REQUIRE(has_prev == 0 || has_prev == 3 || (has_prev == 1 && prev_x), "tz_encode_lossy: bad has_prev / prev_x");
  if (nt == 0) return TZ_OK;
  Geo g = make_geo(H, W, C, Hp, Wp);
  // resident CTAs per SM: threads 32*C, ~13 KB of shared memory -> a persistent grid of a few waves' worth of CTAs
  const int per_sm = 2048 / (32 * C) < 16 ? 2048 / (32 * C) : 16;
  long long grid = (long long)tz::sm_count() * per_sm;
  if (grid > nt) grid = nt;
  cudaStream_t st = (cudaStream_t)stream;
#define TZ_LZ_LAUNCH(CC)                                                                                             \
  lossy_fused_kernel<CC><<<(unsigned)grid, 32 * CC, 0, st>>>(frames, pred_pool, pred_slot, apply, x, nt, g, mode, b0, b1, \
                                                             has_prev, prev_x, hist, overflow, counter)
  if (C == 1) TZ_LZ_LAUNCH(1);
  else if (C == 2) TZ_LZ_LAUNCH(2);
  else if (C == 3) TZ_LZ_LAUNCH(3);
  else TZ_LZ_LAUNCH(4);
#undef TZ_LZ_LAUNCH
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

}  // extern "C"
