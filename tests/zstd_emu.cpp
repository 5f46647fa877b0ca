// Test infrastructure: runs the per-thread bodies of the GPU zstd frame writer (tezip_b200/csrc/tz_zstd_core.h) in
// plain loops on the CPU, in the order of the kernels of tz_zstd.cu, so that the frame layout and the bit streams can
// be checked against libzstd's decoder without a GPU (tests/test_zstd_frames.py compiles this file with g++).  Not
// part of the product: tezip_b200 never loads it.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../tezip_b200/csrc/tz_zstd_core.h"

extern "C" {

unsigned long long emu_bound(unsigned long long n) {
  uint64_t nb = (n + ZS_BLOCK - 1) / ZS_BLOCK;
  return (ZS_FRAME_HDR + n + 3 * (nb + 1) + 8 + 255) & ~255ULL;
}

// zs_hist_kernel
void emu_hist(const uint8_t *src, unsigned long long n, uint32_t *hist, int32_t *uniform) {
  uint64_t nblocks = (n + ZS_BLOCK - 1) / ZS_BLOCK;
  memset(hist, 0, 256 * sizeof(uint32_t));
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint32_t nb = zs_block_len(n, b), h[256] = {0};
    for (uint32_t i = 0; i < nb; ++i) h[src[b * ZS_BLOCK + i]]++;
    int uni = -1;
    for (int s = 0; s < 256; ++s) if (h[s] == nb && nb > 0) uni = s;
    uniform[b] = uni;
    if (uni < 0 && nb >= ZS_MIN_HUF) for (int s = 0; s < 256; ++s) hist[s] += h[s];
  }
}

// tz_zstd_encode: count, scan, layout, prefix, encode.  out: emu_bound(n) bytes, 4-byte aligned.  Returns the size.
unsigned long long emu_encode(const uint8_t *src, unsigned long long n, const uint32_t *ct, const uint8_t *tree,
                              unsigned tree_len, const int32_t *uniform, uint8_t *out) {
  uint64_t nblocks = (n + ZS_BLOCK - 1) / ZS_BLOCK, slots = nblocks * 4 * ZS_SLOTS;
  std::vector<uint16_t> chunk_bits(slots, 0);
  std::vector<uint32_t> chunk_off(slots, 0), stream_bits(nblocks * 4, 0);
  std::vector<ZsBlock> blk(nblocks);
  if (tree_len) {
    for (uint64_t t = 0; t < slots; ++t) {
      uint64_t b = t / (4 * ZS_SLOTS);
      uint32_t s = (uint32_t)(t / ZS_SLOTS) & 3u, j = (uint32_t)(t % ZS_SLOTS), nb = zs_block_len(n, b);
      if (uniform[b] < 0 && nb >= ZS_MIN_HUF)
        chunk_bits[t] = (uint16_t)zs_chunk_bits(src + b * ZS_BLOCK + (uint64_t)s * zs_seg_len(nb, 0),
                                                zs_seg_len(nb, s), j, ct);
    }
    for (uint64_t t = 0; t < nblocks * 4; ++t)
      stream_bits[t] = zs_scan_stream(&chunk_bits[t * ZS_SLOTS], &chunk_off[t * ZS_SLOTS]);
  }
  uint64_t run = ZS_FRAME_HDR;
  for (uint64_t b = 0; b < nblocks; ++b) {
    zs_block_size(zs_block_len(n, b), uniform[b], tree_len, &stream_bits[4 * b], &blk[b]);
    blk[b].off = run;
    run += blk[b].size;
  }
  memset(out, 0, emu_bound(n));
  zs_frame_header(n, out);
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint32_t nb = zs_block_len(n, b);
    uint8_t *dst = out + blk[b].off;
    zs_block_prefix(&blk[b], nb, b + 1 == nblocks, tree, tree_len, dst);
    if (blk[b].type == ZS_RLE) dst[3] = (uint8_t)uniform[b];
    else if (blk[b].type == ZS_RAW) memcpy(dst + 3, src + b * ZS_BLOCK, nb);
    else dst[blk[b].size - 1] = 0;
  }
  if (tree_len) {
    for (uint64_t t = 0; t < slots; ++t) {
      uint64_t b = t / (4 * ZS_SLOTS);
      uint32_t s = (uint32_t)(t / ZS_SLOTS) & 3u, j = (uint32_t)(t % ZS_SLOTS);
      if (blk[b].type != ZS_HUF) continue;
      uint32_t nb = zs_block_len(n, b), seglen = zs_seg_len(nb, s);
      if (j * ZS_CHUNK >= seglen) continue;
      uint64_t byte = blk[b].off + 3 + ZS_LIT_HDR + tree_len + 6;
      for (uint32_t i = 0; i < s; ++i) byte += blk[b].stream_bytes[i];
      zs_encode_chunk(src + b * ZS_BLOCK + (uint64_t)s * zs_seg_len(nb, 0), seglen, j, ct, (uint32_t *)out,
                      byte * 8 + chunk_off[t], (j + 1) * ZS_CHUNK >= seglen);
    }
  }
  return run;
}

// tz_zstd_decode: zs_copy_kernel + zs_decode_kernel.  Returns the largest error code (0 = fine).
int emu_decode(const uint8_t *frame, const ZsDBlock *blk, unsigned long long nblocks, const uint16_t *dtables,
               uint8_t *out) {
  int err = 0;
  for (uint64_t b = 0; b < nblocks; ++b) {
    const ZsDBlock &k = blk[b];
    if (k.type == ZS_RLE) memset(out + k.dst_off, frame[k.src_off], k.regen);
    else if (k.type == ZS_RAW) memcpy(out + k.dst_off, frame + k.src_off, k.regen);
  }
  for (uint64_t t = 0; t < nblocks * 4; ++t) {
    const ZsDBlock &k = blk[t >> 2];
    uint32_t s = (uint32_t)t & 3u;
    if (k.type != ZS_HUF) continue;
    uint64_t so = k.src_off;
    for (uint32_t i = 0; i < s; ++i) so += k.stream_bytes[i];
    int r = zs_decode_stream((const uint32_t *)frame, so, k.stream_bytes[s], dtables + ((uint64_t)k.table << ZS_DLOG),
                             out + k.dst_off + (uint64_t)s * zs_seg_len(k.regen, 0), zs_seg_len(k.regen, s));
    if (r > err) err = r;
  }
  return err;
}

}  // extern "C"
