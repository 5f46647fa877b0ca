// tezip_b200 -- per-thread bodies of the GPU zstd frame writer (tz_zstd.cu).
//
// The container's back-end (compress.py:276,398: zstd.compress(bytes, 9)) only has to produce ONE zstd frame with
// its content size that the reference's zstd.decompress (decompress.py:89,98) reads.  The frames written here use the
// subset of the format (RFC 8878) that has no sequential dependency between blocks: every 128 KB block is an RLE block
// (one repeated byte: the zero frames of the key plane), a raw block, or a compressed block whose literals section
// holds the whole block Huffman-coded in four streams (with its own tree description) and whose sequences section is
// empty.  Every function below is the work of ONE thread on its own piece with offsets that were computed before, so
// that tests/zstd_emu.cpp can run the very same bodies in plain loops on the CPU (no GPU in the build container) and
// check the bytes against libzstd's decoder.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ZS_HD __host__ __device__ __forceinline__
#else
#define ZS_HD inline
#endif

#define ZS_BLOCK 131072u     // Block_Maximum_Size (RFC 8878 3.1.1.2.4): window >= 128 KB
#define ZS_CHUNK 128u        // symbols per encoding thread
#define ZS_SLOTS 256u        // chunk slots per Huffman stream: (ZS_BLOCK / 4) / ZS_CHUNK
#define ZS_FRAME_HDR 14u     // magic 4 + descriptor 1 + window 1 + content size 8
#define ZS_MIN_HUF 1024u     // shorter blocks (only the last one of a frame can be) are stored raw
#define ZS_MAX_TREE 160u     // upper bound of a tree description (1 + 128 direct, FSE-compressed < 128)
#define ZS_LIT_HDR 5u        // literals section header, size format 3 (two 18-bit sizes)

enum { ZS_RAW = 0, ZS_RLE = 1, ZS_HUF = 2 };

struct ZsBlock {             // layout of one block, filled by zs_block_size + the offset scan
  uint64_t off;              // byte offset of the 3-byte block header inside the frame
  uint32_t size;             // bytes of the block including its header
  uint32_t type;             // ZS_RAW / ZS_RLE / ZS_HUF
  uint32_t stream_bytes[4];  // ZS_HUF: bytes of the four Huffman streams
};

ZS_HD uint32_t zs_block_len(uint64_t n, uint64_t b) {
  uint64_t r = n - b * ZS_BLOCK;
  return r < ZS_BLOCK ? (uint32_t)r : ZS_BLOCK;
}
// the four streams of a block of nb literals: three of ceil(nb / 4), the last one takes the rest (RFC 8878 3.1.1.3.1.6)
ZS_HD uint32_t zs_seg_len(uint32_t nb, uint32_t s) {
  uint32_t q = (nb + 3) / 4;
  return s < 3 ? q : nb - 3 * q;
}

struct alignas(16) ZsVec { uint32_t w[4]; };   // 16 bytes of the source (chunks of full blocks start on 128-byte boundaries)

// a full chunk whose first byte is 16-byte aligned is read with eight 16-byte loads, all issued before the first use
ZS_HD bool zs_chunk_vec(const uint8_t *first, uint32_t count) {
  return count == ZS_CHUNK && ((uintptr_t)first & 15) == 0;
}

// Bits of chunk j of a stream.  A stream is written from its LAST symbol to its first (the decoder reads the bit
// stream backwards), so in write order r = 0 .. seglen-1 stands for symbol seglen-1-r; chunk j owns r in
// [j*ZS_CHUNK, (j+1)*ZS_CHUNK).  ct[sym] = code | nbits << 16.
ZS_HD uint32_t zs_chunk_bits(const uint8_t *seg, uint32_t seglen, uint32_t j, const uint32_t *ct) {
  uint32_t lo = j * ZS_CHUNK;
  if (lo >= seglen) return 0;
  uint32_t hi = lo + ZS_CHUNK < seglen ? lo + ZS_CHUNK : seglen;
  uint32_t bits = 0;
  if (zs_chunk_vec(seg + (seglen - hi), hi - lo)) {
    const ZsVec *q = reinterpret_cast<const ZsVec *>(seg + (seglen - hi));
    ZsVec v[ZS_CHUNK / 16];
#pragma unroll
    for (int g = 0; g < (int)(ZS_CHUNK / 16); ++g) v[g] = q[g];
#pragma unroll
    for (int g = 0; g < (int)(ZS_CHUNK / 16); ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t x = v[g].w[i];
        bits += (ct[x & 0xFFu] >> 16) + (ct[(x >> 8) & 0xFFu] >> 16) + (ct[(x >> 16) & 0xFFu] >> 16) + (ct[x >> 24] >> 16);
      }
    return bits;
  }
  for (uint32_t k = seglen - hi; k < seglen - lo; ++k) bits += ct[seg[k]] >> 16;
  return bits;
}

#ifdef __CUDA_ARCH__
#define ZS_OR32(p, v) atomicOr((unsigned int *)(p), (unsigned int)(v))
#else
#define ZS_OR32(p, v) (*(p) |= (uint32_t)(v))
#endif

// Writes chunk j of a stream at absolute bit position `bitpos` of the frame buffer `out` (u32 words, little endian,
// zeroed before).  Words that the chunk shares with its neighbours (its first and its last) are OR-ed atomically, the
// words in between are owned by this chunk alone.  closes: the chunk that holds symbol 0 appends the end mark, the
// single 1 bit above which the decoder starts (RFC 8878 4.2.2).
ZS_HD void zs_encode_chunk(const uint8_t *seg, uint32_t seglen, uint32_t j, const uint32_t *ct, uint32_t *out,
                           uint64_t bitpos, bool closes) {
  uint32_t lo = j * ZS_CHUNK;
  if (lo >= seglen) return;
  uint32_t hi = lo + ZS_CHUNK < seglen ? lo + ZS_CHUNK : seglen;
  uint64_t w = bitpos >> 5;
  uint32_t fill = (uint32_t)(bitpos & 31);
  uint64_t acc = 0;
  bool first = true;
#define ZS_PUT(sym)                                                          \
  do {                                                                       \
    uint32_t e_ = ct[sym];                                                   \
    acc |= (uint64_t)(e_ & 0xFFFFu) << fill;                                 \
    fill += e_ >> 16;                                                        \
    if (fill >= 32) {                                                        \
      if (first) { ZS_OR32(out + w, (uint32_t)acc); first = false; }         \
      else out[w] = (uint32_t)acc;                                           \
      ++w;                                                                   \
      acc >>= 32;                                                            \
      fill -= 32;                                                            \
    }                                                                        \
  } while (0)
  if (zs_chunk_vec(seg + (seglen - hi), hi - lo)) {
    const ZsVec *q = reinterpret_cast<const ZsVec *>(seg + (seglen - hi));
    ZsVec v[ZS_CHUNK / 16];
#pragma unroll
    for (int g = 0; g < (int)(ZS_CHUNK / 16); ++g) v[g] = q[g];
#pragma unroll
    for (int g = (int)(ZS_CHUNK / 16) - 1; g >= 0; --g)
#pragma unroll
      for (int i = 3; i >= 0; --i) {
        uint32_t x = v[g].w[i];
        ZS_PUT(x >> 24); ZS_PUT((x >> 16) & 0xFFu); ZS_PUT((x >> 8) & 0xFFu); ZS_PUT(x & 0xFFu);
      }
  } else {
    for (uint32_t k = seglen - lo; k-- > seglen - hi;) ZS_PUT(seg[k]);
  }
#undef ZS_PUT
  if (closes) {
    acc |= (uint64_t)1 << fill;
    ++fill;
    if (fill >= 32) {
      if (first) { ZS_OR32(out + w, (uint32_t)acc); first = false; }
      else out[w] = (uint32_t)acc;
      ++w;
      acc >>= 32;
      fill -= 32;
    }
  }
  if (fill) ZS_OR32(out + w, (uint32_t)acc);
}

// Exclusive scan of the chunk bit counts of one stream (ZS_SLOTS slots) -> bit offset of every chunk inside the
// stream; returns the stream's bits.
ZS_HD uint32_t zs_scan_stream(const uint16_t *chunk_bits, uint32_t *chunk_off) {
  uint32_t run = 0;
  for (uint32_t j = 0; j < ZS_SLOTS; ++j) {
    chunk_off[j] = run;
    run += chunk_bits[j];
  }
  return run;
}

// Type and size of block b.  uniform >= 0: all bytes of the block equal that value.  stream_bits: the four streams'
// bit counts (only read when a Huffman table exists: tree_len > 0).
ZS_HD void zs_block_size(uint32_t nb, int uniform, uint32_t tree_len, const uint32_t *stream_bits, ZsBlock *blk) {
  blk->stream_bytes[0] = blk->stream_bytes[1] = blk->stream_bytes[2] = blk->stream_bytes[3] = 0;
  if (uniform >= 0 && nb > 0) {
    blk->type = ZS_RLE;
    blk->size = 3 + 1;
    return;
  }
  if (tree_len > 0 && nb >= ZS_MIN_HUF) {
    uint32_t c = 3 + ZS_LIT_HDR + tree_len + 6 + 1;     // block header, literals header, tree, jump table, "0 sequences"
    for (int s = 0; s < 4; ++s) {
      blk->stream_bytes[s] = (stream_bits[s] >> 3) + 1;  // + the end mark bit, rounded up to bytes
      c += blk->stream_bytes[s];
    }
    if (c < 3 + nb) {
      blk->type = ZS_HUF;
      blk->size = c;
      return;
    }
  }
  blk->type = ZS_RAW;
  blk->size = 3 + nb;
}

// The bytes in front of a block's payload: block header (3 bytes: last flag, type, size) and, for a Huffman block,
// literals section header, tree description and jump table.  Returns the number of bytes written to `dst`
// (<= 3 + 5 + ZS_MAX_TREE + 6).  The caller stores the payload behind them: raw bytes, the RLE byte, or the four
// streams followed by one zero byte (Number_of_Sequences = 0).
ZS_HD uint32_t zs_block_prefix(const ZsBlock *blk, uint32_t nb, bool last, const uint8_t *tree, uint32_t tree_len,
                               uint8_t *dst) {
  uint32_t content = blk->type == ZS_RLE ? nb : blk->size - 3;   // RLE: Block_Size is the regenerated size
  uint32_t h = (last ? 1u : 0u) | (blk->type << 1) | (content << 3);
  dst[0] = (uint8_t)h; dst[1] = (uint8_t)(h >> 8); dst[2] = (uint8_t)(h >> 16);
  if (blk->type != ZS_HUF) return 3;
  uint32_t csize = tree_len + 6 + blk->stream_bytes[0] + blk->stream_bytes[1] + blk->stream_bytes[2] +
                   blk->stream_bytes[3];
  uint64_t lh = 2u | (3u << 2) | ((uint64_t)nb << 4) | ((uint64_t)csize << 22);   // Compressed_Literals_Block, format 3
  for (int i = 0; i < 5; ++i) dst[3 + i] = (uint8_t)(lh >> (8 * i));
  uint32_t o = 8;
  for (uint32_t i = 0; i < tree_len; ++i) dst[o++] = tree[i];
  for (int s = 0; s < 3; ++s) {
    dst[o++] = (uint8_t)blk->stream_bytes[s];
    dst[o++] = (uint8_t)(blk->stream_bytes[s] >> 8);
  }
  return o;
}

// Frame header: magic, descriptor 0xC0 (8-byte content size, no single-segment flag, no checksum, no dictionary),
// window descriptor 0x38 (2^17 bytes = one block: nothing ever refers back), content size.
ZS_HD void zs_frame_header(uint64_t n, uint8_t *dst) {
  dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
  dst[4] = 0xC0;
  dst[5] = 0x38;
  for (int i = 0; i < 8; ++i) dst[6 + i] = (uint8_t)(n >> (8 * i));
}

// ---------------------------------------------------------------------------------------------------------------------
// Reading such frames back (decompress.py:89,98): the host walks the block headers (tezip_b200/zstd_frames.py), the
// device decodes every block -- and every one of the four Huffman streams of a block -- independently.

struct ZsDBlock {            // one block of a parsed frame
  uint64_t src_off;          // frame offset of the payload: raw bytes, the RLE byte, or the first Huffman stream
  uint64_t dst_off;          // offset of the block's bytes in the decoded content
  uint32_t type;             // ZS_RAW / ZS_RLE / ZS_HUF
  uint32_t regen;            // decoded bytes of the block
  uint32_t stream_bytes[4];  // ZS_HUF: bytes of the four streams
  uint32_t table;            // ZS_HUF: index of its decoding table
  uint32_t pad;
};

#define ZS_DLOG 11u          // decoding tables are expanded to 2^11 entries: symbol | nbits << 8

// Decodes one Huffman stream of nsym symbols (RFC 8878 4.2.2): the stream is read from its last byte down, starting
// below the highest set bit of that byte.  The frame is read as aligned 32-bit words (frame: 4-byte aligned base, the
// stream occupies bytes [off, off + nbytes)); a refill may load words that lie below the stream's first byte, but a
// well-formed stream never consumes their bits: `left` counts the stream's own bits and must end at exactly zero.
// Four symbols are assembled per 32-bit store where dst allows.  Returns 0, or a non-zero code when the stream is
// malformed (no end mark, more bits consumed than it holds, bits left over).
ZS_HD int zs_decode_stream(const uint32_t *frame, uint64_t off, uint32_t nbytes, const uint16_t *dtable, uint8_t *dst,
                           uint32_t nsym) {
  if (nbytes == 0) return 1;
  const uint32_t last = reinterpret_cast<const uint8_t *>(frame)[off + nbytes - 1];
  if (last == 0) return 2;
  int top = 7;
  while (!((last >> top) & 1)) --top;                   // the end mark; `top` payload bits lie below it in this byte
  const uint64_t mark = (off + nbytes - 1) * 8 + (uint32_t)top;   // absolute bit position of the end mark
  long long left = (long long)(nbytes - 1) * 8 + top;             // payload bits of the stream
  long long wi = (long long)(mark >> 5);
  const uint32_t b = (uint32_t)mark & 31u;
  uint64_t acc = b ? (uint64_t)(frame[wi] & ((1u << b) - 1u)) << (64 - b) : 0;   // next bit to read = bit 63
  int have = (int)b;                                                               // bits loaded into acc
  --wi;
#define ZS_REFILL()                                     \
  if (have <= 32) {                                     \
    uint32_t w_ = wi >= 0 ? frame[wi] : 0u;             \
    --wi;                                               \
    acc |= (uint64_t)w_ << (32 - have);                 \
    have += 32;                                         \
  }
#define ZS_DEC(v)                                       \
  {                                                     \
    uint32_t e_ = dtable[acc >> (64 - ZS_DLOG)];        \
    uint32_t nb_ = e_ >> 8;                             \
    (v) = e_ & 0xFFu;                                   \
    acc <<= nb_;                                        \
    have -= (int)nb_;                                   \
    left -= nb_;                                        \
  }
  uint32_t i = 0, v0, v1, v2, v3;
  while (i < nsym && ((uintptr_t)(dst + i) & 3)) {      // head: up to the first aligned output word
    ZS_REFILL();
    ZS_DEC(v0);
    dst[i++] = (uint8_t)v0;
  }
  for (; i + 4 <= nsym; i += 4) {
    ZS_REFILL();
    ZS_DEC(v0); ZS_DEC(v1);
    ZS_REFILL();
    ZS_DEC(v2); ZS_DEC(v3);
    *reinterpret_cast<uint32_t *>(dst + i) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
  }
  for (; i < nsym; ++i) {
    ZS_REFILL();
    ZS_DEC(v0);
    dst[i] = (uint8_t)v0;
  }
#undef ZS_REFILL
#undef ZS_DEC
  return left == 0 ? 0 : (left < 0 ? 3 : 4);
}
