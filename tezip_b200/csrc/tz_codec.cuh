// tezip_b200 -- pieces shared by the 8-bit codec kernels (tz_codec.cu) and the 16-bit / wide-code ones
// (tz_codec_wide.cu): frame geometry, the warp-serial error_bound scan, the chunk-sum scan, grid sizing.
#pragma once
#include "tz_common.cuh"

namespace {

struct Geo {
  int H, W, C, Hp, Wp;
  int rowlen;              // W*C   samples per cropped row
  int prow;                // Wp*C  floats per padded row
  long long frame_elems;   // H*W*C
  long long pframe_elems;  // Hp*Wp*C
};

static Geo make_geo(int H, int W, int C, int Hp, int Wp) {
  Geo g;
  g.H = H; g.W = W; g.C = C; g.Hp = Hp; g.Wp = Wp;
  g.rowlen = W * C;
  g.prow = Wp * C;
  g.frame_elems = (long long)H * W * C;
  g.pframe_elems = (long long)Hp * Wp * C;
  return g;
}

// compress.py:307,310-311: float32 product, then truncation toward zero.
__device__ __forceinline__ int q255(float p) { return __float2int_rz(__fmul_rn(p, 255.0f)); }

// i - f * d where 0 <= i - f * d < 2^31 (the offset of stream element i inside its frame), in UNSIGNED 32-bit
// arithmetic.  Written as `(int)(i - f * d)` nvcc narrows the expression to 32-bit operations on the low words and
// then widens the PARTS as if they were non-overflowing signed ints; when the low words of i and f * d lie on
// different sides of 2^31 the pool address is off by 2^32 elements.  Seen at stream element 2^31 - 1 of a
// 1024x1024x1 sequence (frame 2048): one wrong delta, and every later frame decoded with a constant offset.
__device__ __forceinline__ int rem_in_frame(long long i, long long f, long long d) {
  const unsigned int lo = (unsigned int)(unsigned long long)i;
  const unsigned int fd = (unsigned int)((unsigned long long)f * (unsigned long long)d);
  return (int)(lo - fd);
}

// compress.py:293-314 for one sample (generic addressing).
__device__ __forceinline__ int resid_at(const uint8_t *__restrict__ frames, const float *__restrict__ pool,
                                        const int32_t *__restrict__ slot, const Geo &g, long long i) {
  long long f = i / g.frame_elems;
  int s = slot[f];
  if (s < 0) return 0;
  int r = rem_in_frame(i, f, g.frame_elems);
  int row = r / g.rowlen;
  int col = r - row * g.rowlen;
  float p = pool[(long long)s * g.pframe_elems + (long long)row * g.prow + col];
  return q255(p) - (int)frames[i];
}

// Loads the residuals of 8 consecutive samples i0..i0+7 (i0 % 8 == 0) into v[0..7].
// Fast path (rowlen % 8 == 0): one 8-byte frame load + two 16-byte prediction loads.
template <bool FAST>
__device__ __forceinline__ void resid8(const uint8_t *__restrict__ frames, const float *__restrict__ pool,
                                       const int32_t *__restrict__ slot, const Geo &g, long long i0,
                                       long long n, int v[8]) {
  if (FAST) {
    long long f = i0 / g.frame_elems;
    int s = slot[f];
    if (s < 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = 0;
      return;
    }
    int r = rem_in_frame(i0, f, g.frame_elems);
    int row = r / g.rowlen;
    int col = r - row * g.rowlen;
    uint2 a = *reinterpret_cast<const uint2 *>(frames + i0);
    const float4 *pp =
        reinterpret_cast<const float4 *>(pool + (long long)s * g.pframe_elems + (long long)row * g.prow + col);
    float4 p0 = pp[0], p1 = pp[1];
    v[0] = q255(p0.x) - (int)(a.x & 0xff);
    v[1] = q255(p0.y) - (int)((a.x >> 8) & 0xff);
    v[2] = q255(p0.z) - (int)((a.x >> 16) & 0xff);
    v[3] = q255(p0.w) - (int)(a.x >> 24);
    v[4] = q255(p1.x) - (int)(a.y & 0xff);
    v[5] = q255(p1.y) - (int)((a.y >> 8) & 0xff);
    v[6] = q255(p1.z) - (int)((a.y >> 16) & 0xff);
    v[7] = q255(p1.w) - (int)(a.y >> 24);
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (i0 + k < n) ? resid_at(frames, pool, slot, g, i0 + k) : 0;
  }
}

__device__ __forceinline__ void load8_i16(const int16_t *__restrict__ x, long long i0, long long n, int v[8]) {
  if (i0 + 8 <= n) {
    uint4 a = *reinterpret_cast<const uint4 *>(x + i0);
    v[0] = (int16_t)(a.x & 0xffff); v[1] = (int16_t)(a.x >> 16);
    v[2] = (int16_t)(a.y & 0xffff); v[3] = (int16_t)(a.y >> 16);
    v[4] = (int16_t)(a.z & 0xffff); v[5] = (int16_t)(a.z >> 16);
    v[6] = (int16_t)(a.w & 0xffff); v[7] = (int16_t)(a.w >> 16);
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (i0 + k < n) ? (int)x[i0 + k] : 0;
  }
}

__device__ __forceinline__ void store8_i16(int16_t *__restrict__ out, long long i0, long long n, const int v[8]) {
  if (i0 + 8 <= n) {
    uint4 a;
    a.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
    a.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
    a.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
    a.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
    *reinterpret_cast<uint4 *>(out + i0) = a;
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (i0 + k < n) out[i0 + k] = (int16_t)v[k];
  }
}

// compress.py:75: y[i] = x[i-1] - x[i] in int16 arithmetic; y[0] = x[0] (or prev_x - x[0] on a shard).
__device__ __forceinline__ void delta8(const int v[8], int prev, bool is_first_global, int y[8]) {
  y[0] = is_first_global ? v[0] : (int)(int16_t)(prev - v[0]);
#pragma unroll
  for (int k = 1; k < 8; k++) y[k] = (int)(int16_t)(v[k - 1] - v[k]);
}

struct EbInt {   // running (min d, max d)
  int a, b;
  __device__ static EbInt empty() { return {2147483647, -2147483647 - 1}; }
  __device__ static EbInt of(int d, double, bool in) { return in ? EbInt{d, d} : empty(); }
  __device__ EbInt join(const EbInt &o) const { return {min(a, o.a), max(b, o.b)}; }
  __device__ bool broken(int G) const { return (long long)b - (long long)a > (long long)G; }
  __device__ double mid(double E) const {   // (fl(dmin+E) + fl(dmax-E)) / 2, compress.py:61
    return __dmul_rn(__dadd_rn(__dadd_rn((double)a, E), __dsub_rn((double)b, E)), 0.5);
  }
  __device__ EbInt shfl(int src) const { return {__shfl_sync(0xffffffffu, a, src), __shfl_sync(0xffffffffu, b, src)}; }
  __device__ EbInt shfl_up(int o) const { return {__shfl_up_sync(0xffffffffu, a, o), __shfl_up_sync(0xffffffffu, b, o)}; }
};
struct EbDbl {   // running (u = min Du, l = max Dl)
  double u, l;
  __device__ static EbDbl empty() {
    return {__longlong_as_double(0x7ff0000000000000LL), __longlong_as_double(0xfff0000000000000LL)};
  }
  __device__ static EbDbl of(int d, double e, bool in) {
    return in ? EbDbl{__dadd_rn((double)d, e), __dsub_rn((double)d, e)} : empty();   // compress.py:47-48
  }
  __device__ EbDbl join(const EbDbl &o) const { return {(o.u < u) ? o.u : u, (o.l > l) ? o.l : l}; }
  __device__ bool broken(int) const { return __dsub_rn(u, l) < 0.0; }                 // compress.py:60
  __device__ double mid(double) const { return __dmul_rn(__dadd_rn(u, l), 0.5); }     // compress.py:61
  __device__ EbDbl shfl(int src) const { return {__shfl_sync(0xffffffffu, u, src), __shfl_sync(0xffffffffu, l, src)}; }
  __device__ EbDbl shfl_up(int o) const { return {__shfl_up_sync(0xffffffffu, u, o), __shfl_up_sync(0xffffffffu, l, o)}; }
};

template <typename VAL, typename PIX = uint8_t, typename CODE = int16_t>
__device__ __forceinline__ void eb_plane_warp(const PIX *__restrict__ o, CODE *__restrict__ d, int n, int C,
                                              bool pwrel, double b0, double E, int G) {
  const int lane = threadIdx.x & 31;
  VAL carry = VAL::empty();   // state of the open segment [head, ...)
  int head = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool in = i < n;
    const int dv = in ? (int)d[(long long)i * C] : 0;
    const double e = pwrel ? __dmul_rn((double)(in ? (int)o[(long long)i * C] : 0), b0) : E;   // compress.py:45
    const VAL mine = VAL::of(dv, e, in);
    // ---- A: where does the carried segment end?
    VAL pre = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      VAL t = pre.shfl_up(off);
      if (lane >= off) pre = pre.join(t);
    }
    pre = pre.join(carry);
    const unsigned m = __ballot_sync(0xffffffffu, in && pre.broken(G));
    if (m == 0) {   // the whole chunk joins the open segment
      carry = pre.shfl(31);
      continue;
    }
    const int b = __ffs(m) - 1;
    {
      const VAL seg = (b > 0) ? pre.shfl(b - 1) : carry;
      const CODE q = (CODE)(long long)seg.mid(E);               // float -> int64 slice assignment truncates
      for (int j = head + lane; j < base; j += 32) d[(long long)j * C] = q;   // part of the segment behind this chunk
      if (lane < b && head <= i) d[(long long)i * C] = q;
    }
    // ---- B: for a segment starting at lane s, the first lane that breaks it (32 = still open at chunk end)
    VAL run = mine;
    int nb = 32;
    bool done = !in;
    for (int k = 1; k < 32; k++) {
      const int t = lane + k;
      const VAL vt = mine.shfl(t & 31);
      const int tin_i = __shfl_sync(0xffffffffu, (int)in, t & 31);   // executed by every lane: no short-circuit
      const bool tin = (t < 32) && (tin_i != 0);
      if (!done) {
        if (!tin) {
          done = true;
        } else {
          const VAL nx = run.join(vt);
          if (nx.broken(G)) {
            done = true;
            nb = t;
          } else {
            run = nx;
          }
        }
      }
      if (__all_sync(0xffffffffu, done)) break;
    }
    // ---- C: chase the links from lane b
    int cur = b;
    CODE myq = 0;
    bool have = false;
    for (int guard = 0;; guard++) {
      if (guard > 40) {   // cannot happen (links strictly increase); never hang the GPU on a logic error
        if (lane == 0) printf("tezip_b200: error_bound link chase did not terminate (base %d cur %d)\n", base, cur);
        __trap();
      }
      const int nxt = __shfl_sync(0xffffffffu, nb, cur);
      const VAL seg = run.shfl(cur);
      if (nxt >= 32) {   // stays open: becomes the carried segment
        carry = seg;
        head = base + cur;
        break;
      }
      const CODE q = (CODE)(long long)seg.mid(E);
      if (lane >= cur && lane < nxt) {
        myq = q;
        have = true;
      }
      cur = nxt;
    }
    if (have) d[(long long)i * C] = myq;
  }
  if (head < n) {   // compress.py:67
    const CODE q = (CODE)(long long)carry.mid(E);
    for (int j = head + lane; j < n; j += 32) d[(long long)j * C] = q;
  }
}

constexpr int DEC_THREADS = 256;
constexpr int DEC_CHUNK = DEC_THREADS * 8;

// exclusive scan of the chunk sums, single block (at most a few 10^4 chunks).
__global__ void __launch_bounds__(1024) decode_scan_kernel(unsigned int *__restrict__ sums, long long nchunks) {
  __shared__ unsigned int wtot[32];
  __shared__ unsigned int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < nchunks; base += 1024) {
    long long i = base + threadIdx.x;
    unsigned int v = (i < nchunks) ? sums[i] : 0;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned int w = wtot[threadIdx.x];
      unsigned int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (threadIdx.x >= o) winc += t;
      }
      wtot[threadIdx.x] = winc - w;   // exclusive warp offsets
    }
    __syncthreads();
    unsigned int carry = carry_s;
    unsigned int excl = carry + wtot[threadIdx.x >> 5] + inc - v;
    if (i < nchunks) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wtot[31] + inc;
    __syncthreads();
  }
}

static int stream_grid(long long work_items, int threads, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)tz::sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static bool fast_ok(const Geo &g) { return (g.rowlen % 8) == 0; }

}  // namespace
