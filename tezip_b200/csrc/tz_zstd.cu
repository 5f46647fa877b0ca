// tezip_b200 -- GPU writer of zstd frames for the container's back-end (compress.py:276,398).
// See tz_zstd_core.h for the subset of the format and the per-thread bodies; this file holds the kernels around them
// and the C ABI.  HBM-bound byte work: three reads of the source (histogram, bit counts, encode) and one write.
#include "tz_common.cuh"
#include "tz_zstd_core.h"

namespace {

// One CTA per 128 KB block: byte histogram in shared memory; uniform[b] = the byte if the block holds one value
// only (it becomes an RLE block), else -1 and the counts are added to the frame's histogram (blocks shorter than
// ZS_MIN_HUF are stored raw and stay out of it).
__global__ void __launch_bounds__(512) zs_hist_kernel(const uint8_t *__restrict__ src, uint64_t n,
                                                      uint32_t *__restrict__ hist, int32_t *__restrict__ uniform) {
  __shared__ uint32_t h[256];
  __shared__ int uni;
  const uint64_t b = blockIdx.x;
  const uint32_t nb = zs_block_len(n, b);
  const uint8_t *p = src + b * ZS_BLOCK;
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  if (threadIdx.x == 0) uni = -1;
  __syncthreads();
  if (nb == ZS_BLOCK && ((uintptr_t)p & 15) == 0) {   // full blocks of an aligned source: 16-byte loads
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    for (uint32_t i = threadIdx.x; i < ZS_BLOCK / 16; i += blockDim.x) {
      uint4 v = q[i];
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // a word of four equal bytes (zero frames of the key plane, constant high bytes) costs one atomic
        uint32_t x = w[k];
        if (x == (x & 0xFFu) * 0x01010101u) atomicAdd(&h[x & 0xFFu], 4u);
        else {
          atomicAdd(&h[x & 0xFFu], 1u); atomicAdd(&h[(x >> 8) & 0xFFu], 1u);
          atomicAdd(&h[(x >> 16) & 0xFFu], 1u); atomicAdd(&h[x >> 24], 1u);
        }
      }
    }
  } else {
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) atomicAdd(&h[p[i]], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256 && h[threadIdx.x] == nb && nb > 0) uni = (int)threadIdx.x;
  __syncthreads();
  if (threadIdx.x == 0) uniform[b] = uni;
  if (uni < 0 && nb >= ZS_MIN_HUF && threadIdx.x < 256 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// One thread per chunk slot (block, stream, slot): bit count of the chunk.
__global__ void __launch_bounds__(256) zs_count_kernel(const uint8_t *__restrict__ src, uint64_t n,
                                                       const uint32_t *__restrict__ ct_g,
                                                       const int32_t *__restrict__ uniform,
                                                       uint16_t *__restrict__ chunk_bits, uint64_t slots) {
  __shared__ uint32_t ct[256];
  ct[threadIdx.x] = ct_g[threadIdx.x];
  __syncthreads();
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= slots) return;
  const uint64_t b = t / (4 * ZS_SLOTS);
  const uint32_t s = (uint32_t)(t / ZS_SLOTS) & 3u, j = (uint32_t)(t % ZS_SLOTS);
  const uint32_t nb = zs_block_len(n, b);
  uint32_t bits = 0;
  if (uniform[b] < 0 && nb >= ZS_MIN_HUF)
    bits = zs_chunk_bits(src + b * ZS_BLOCK + (uint64_t)s * zs_seg_len(nb, 0), zs_seg_len(nb, s), j, ct);
  chunk_bits[t] = (uint16_t)bits;
}

// One thread per stream: offsets of its chunks, its bit count.
__global__ void zs_scan_kernel(const uint16_t *__restrict__ chunk_bits, uint32_t *__restrict__ chunk_off,
                               uint32_t *__restrict__ stream_bits, uint64_t streams) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= streams) return;
  stream_bits[t] = zs_scan_stream(chunk_bits + t * ZS_SLOTS, chunk_off + t * ZS_SLOTS);
}

// One CTA: types and sizes of all blocks, then their offsets (each thread scans a contiguous range of blocks, thread 0
// scans the range totals).  total[0] = bytes of the frame.
__global__ void __launch_bounds__(1024) zs_layout_kernel(uint64_t n, uint64_t nblocks, const int32_t *__restrict__ uniform,
                                                         uint32_t tree_len, const uint32_t *__restrict__ stream_bits,
                                                         ZsBlock *__restrict__ blk, uint64_t *__restrict__ total) {
  __shared__ uint64_t part[1024];
  const uint64_t per = (nblocks + blockDim.x - 1) / blockDim.x;
  const uint64_t b0 = threadIdx.x * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
  uint64_t sum = 0;
  for (uint64_t b = b0; b < b1; ++b) {
    zs_block_size(zs_block_len(n, b), uniform[b], tree_len, stream_bits + 4 * b, blk + b);
    sum += blk[b].size;
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t run = ZS_FRAME_HDR;
    for (uint32_t i = 0; i < blockDim.x; ++i) {
      uint64_t v = part[i];
      part[i] = run;
      run += v;
    }
    total[0] = run;
  }
  __syncthreads();
  uint64_t run = part[threadIdx.x];
  for (uint64_t b = b0; b < b1; ++b) {
    blk[b].off = run;
    run += blk[b].size;
  }
}

// One CTA per block: everything of the block except the Huffman streams (frame header by block 0).
__global__ void __launch_bounds__(256) zs_prefix_kernel(const uint8_t *__restrict__ src, uint64_t n, uint64_t nblocks,
                                                        const int32_t *__restrict__ uniform,
                                                        const uint8_t *__restrict__ tree, uint32_t tree_len,
                                                        const ZsBlock *__restrict__ blk, uint8_t *__restrict__ out) {
  __shared__ uint8_t pre[3 + ZS_LIT_HDR + ZS_MAX_TREE + 6];
  __shared__ uint32_t plen;
  const uint64_t b = blockIdx.x;
  const uint32_t nb = zs_block_len(n, b);
  const ZsBlock k = blk[b];
  if (threadIdx.x == 0) {
    plen = zs_block_prefix(&k, nb, b + 1 == nblocks, tree, tree_len, pre);
    if (b == 0) zs_frame_header(n, out);
  }
  __syncthreads();
  uint8_t *dst = out + k.off;
  for (uint32_t i = threadIdx.x; i < plen; i += blockDim.x) dst[i] = pre[i];
  if (k.type == ZS_RLE) {
    if (threadIdx.x == 0) dst[3] = (uint8_t)uniform[b];
  } else if (k.type == ZS_RAW) {
    const uint8_t *p = src + b * ZS_BLOCK;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) dst[3 + i] = p[i];
  } else if (threadIdx.x == 0) {
    dst[k.size - 1] = 0;                    // sequences section: Number_of_Sequences = 0
  }
}

// One thread per chunk slot: the chunk's codes at their final bit position.
__global__ void __launch_bounds__(256) zs_encode_kernel(const uint8_t *__restrict__ src, uint64_t n,
                                                        const uint32_t *__restrict__ ct_g, uint32_t tree_len,
                                                        const ZsBlock *__restrict__ blk,
                                                        const uint32_t *__restrict__ chunk_off,
                                                        uint32_t *__restrict__ out, uint64_t slots) {
  __shared__ uint32_t ct[256];
  ct[threadIdx.x] = ct_g[threadIdx.x];
  __syncthreads();
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= slots) return;
  const uint64_t b = t / (4 * ZS_SLOTS);
  const uint32_t s = (uint32_t)(t / ZS_SLOTS) & 3u, j = (uint32_t)(t % ZS_SLOTS);
  if (blk[b].type != ZS_HUF) return;
  const uint32_t nb = zs_block_len(n, b), seglen = zs_seg_len(nb, s);
  if (j * ZS_CHUNK >= seglen) return;
  uint64_t byte = blk[b].off + 3 + ZS_LIT_HDR + tree_len + 6;
  for (uint32_t i = 0; i < s; ++i) byte += blk[b].stream_bytes[i];
  zs_encode_chunk(src + b * ZS_BLOCK + (uint64_t)s * zs_seg_len(nb, 0), seglen, j, ct, out,
                  byte * 8 + chunk_off[t], (j + 1) * ZS_CHUNK >= seglen);
}

// Decoder: one thread per Huffman stream; a CTA (one warp) holds the streams of eight blocks.  The decoding table of
// the CTA's first block is staged in shared memory -- the blocks of one frame nearly always share one table -- and a
// block with another table reads its own through L1.  The kernel is latency-bound by construction (a stream is a
// serial chain), so the per-symbol chain is what counts: one table look-up, one shift.
__global__ void __launch_bounds__(32) zs_decode_kernel(const uint8_t *__restrict__ frame, const ZsDBlock *__restrict__ blk,
                                                       uint64_t nblocks, const uint16_t *__restrict__ dtables,
                                                       uint8_t *__restrict__ out, int32_t *__restrict__ err) {
  __shared__ uint16_t tab[1u << ZS_DLOG];
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b = t >> 2;
  const uint32_t s = (uint32_t)t & 3u;
  const uint32_t tab0 = blk[(uint64_t)blockIdx.x * 8].table;          // (0 for raw / RLE blocks: any valid table)
  {
    const uint4 *g = reinterpret_cast<const uint4 *>(dtables + ((uint64_t)tab0 << ZS_DLOG));
    uint4 *d = reinterpret_cast<uint4 *>(tab);
    for (uint32_t i = threadIdx.x; i < (2u << ZS_DLOG) / 16; i += 32) d[i] = g[i];
  }
  __syncthreads();
  if (b >= nblocks) return;
  const ZsDBlock k = blk[b];
  if (k.type != ZS_HUF) return;
  const uint64_t so = k.src_off + (s > 0 ? k.stream_bytes[0] : 0u) + (s > 1 ? k.stream_bytes[1] : 0u) +
                      (s > 2 ? k.stream_bytes[2] : 0u);
  const uint32_t nbytes = s == 0 ? k.stream_bytes[0] : s == 1 ? k.stream_bytes[1] : s == 2 ? k.stream_bytes[2]
                                                                                             : k.stream_bytes[3];
  const uint32_t q = zs_seg_len(k.regen, 0);
  const uint16_t *table = k.table == tab0 ? tab : dtables + ((uint64_t)k.table << ZS_DLOG);
  int r = zs_decode_stream(reinterpret_cast<const uint32_t *>(frame), so, nbytes, table,
                           out + k.dst_off + (uint64_t)s * q, zs_seg_len(k.regen, s));
  if (r) atomicMax(err, r);
}

// Raw and RLE blocks: one CTA per block.
__global__ void __launch_bounds__(256) zs_copy_kernel(const uint8_t *__restrict__ frame, const ZsDBlock *__restrict__ blk,
                                                      uint8_t *__restrict__ out) {
  const ZsDBlock k = blk[blockIdx.x];
  if (k.type == ZS_HUF) return;
  uint8_t *dst = out + k.dst_off;
  const uint8_t *p = frame + k.src_off;
  if (k.type == ZS_RLE) {
    const uint8_t v = p[0];
    for (uint32_t i = threadIdx.x; i < k.regen; i += blockDim.x) dst[i] = v;
  } else {
    for (uint32_t i = threadIdx.x; i < k.regen; i += blockDim.x) dst[i] = p[i];
  }
}

inline uint64_t n_blocks(uint64_t n) { return (n + ZS_BLOCK - 1) / ZS_BLOCK; }
inline uint64_t align256(uint64_t v) { return (v + 255) & ~(uint64_t)255; }

}  // namespace

extern "C" {

unsigned long long tz_zstd_bound(unsigned long long n) {
  return align256(ZS_FRAME_HDR + n + 3 * (n_blocks(n) + 1) + 8);
}

unsigned long long tz_zstd_workspace_bytes(unsigned long long n) {
  const uint64_t nb = n_blocks(n) + 1;
  return align256(nb * 4 * ZS_SLOTS * sizeof(uint16_t)) + align256(nb * 4 * ZS_SLOTS * sizeof(uint32_t)) +
         align256(nb * 4 * sizeof(uint32_t)) + align256(nb * sizeof(ZsBlock)) + 256;
}

int tz_zstd_hist(const uint8_t *src, unsigned long long n, uint32_t *hist, int32_t *uniform, void *stream) {
  TZ_REQUIRE(src && hist && uniform && n > 0 && n_blocks(n) < 2147483647ULL, "tz_zstd_hist: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  TZ_CHECK_CUDA(cudaMemsetAsync(hist, 0, 256 * sizeof(uint32_t), st));
  zs_hist_kernel<<<(unsigned)n_blocks(n), 512, 0, st>>>(src, n, hist, uniform);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

int tz_zstd_encode(const uint8_t *src, unsigned long long n, const uint32_t *ctable, const uint8_t *tree,
                   unsigned tree_len, const int32_t *uniform, void *workspace, uint8_t *out,
                   unsigned long long *total, void *stream) {
  TZ_REQUIRE(src && ctable && uniform && workspace && out && total && n > 0 && n_blocks(n) < 2147483647ULL &&
             tree_len <= ZS_MAX_TREE && (tree_len == 0 || tree) && ((uintptr_t)out & 3) == 0,
             "tz_zstd_encode: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t nb = n_blocks(n), slots = nb * 4 * ZS_SLOTS, streams = nb * 4;
  uint8_t *w = (uint8_t *)workspace;
  uint16_t *chunk_bits = (uint16_t *)w;  w += align256((nb + 1) * 4 * ZS_SLOTS * sizeof(uint16_t));
  uint32_t *chunk_off = (uint32_t *)w;   w += align256((nb + 1) * 4 * ZS_SLOTS * sizeof(uint32_t));
  uint32_t *stream_bits = (uint32_t *)w; w += align256((nb + 1) * 4 * sizeof(uint32_t));
  ZsBlock *blk = (ZsBlock *)w;
  if (tree_len) {
    zs_count_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(src, n, ctable, uniform, chunk_bits, slots);
    TZ_CHECK_LAUNCH();
    zs_scan_kernel<<<(unsigned)((streams + 127) / 128), 128, 0, st>>>(chunk_bits, chunk_off, stream_bits, streams);
    TZ_CHECK_LAUNCH();
  }
  zs_layout_kernel<<<1, 1024, 0, st>>>(n, nb, uniform, tree_len, stream_bits, blk, (uint64_t *)total);
  TZ_CHECK_LAUNCH();
  TZ_CHECK_CUDA(cudaMemsetAsync(out, 0, tz_zstd_bound(n), st));
  zs_prefix_kernel<<<(unsigned)nb, 256, 0, st>>>(src, n, nb, uniform, tree, tree_len, blk, out);
  TZ_CHECK_LAUNCH();
  if (tree_len) {
    zs_encode_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(src, n, ctable, tree_len, blk, chunk_off,
                                                                      (uint32_t *)out, slots);
    TZ_CHECK_LAUNCH();
  }
  return TZ_OK;
}

int tz_zstd_decode(const uint8_t *frame, const void *blocks, unsigned long long nblocks, const uint16_t *dtables,
                   uint8_t *out, int32_t *err, void *stream) {
  TZ_REQUIRE(frame && blocks && out && err && nblocks > 0 && nblocks < 2147483647ULL / 4 && ((uintptr_t)frame & 3) == 0 &&
             ((uintptr_t)dtables & 15) == 0, "tz_zstd_decode: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  TZ_CHECK_CUDA(cudaMemsetAsync(err, 0, sizeof(int32_t), st));
  zs_copy_kernel<<<(unsigned)nblocks, 256, 0, st>>>(frame, (const ZsDBlock *)blocks, out);
  TZ_CHECK_LAUNCH();
  if (dtables) {
    zs_decode_kernel<<<(unsigned)((nblocks * 4 + 31) / 32), 32, 0, st>>>(frame, (const ZsDBlock *)blocks, nblocks, dtables,
                                                                        out, err);
    TZ_CHECK_LAUNCH();
  }
  return TZ_OK;
}

}  // extern "C"
