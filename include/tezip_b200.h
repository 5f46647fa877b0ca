/* tezip_b200 -- C ABI of the B200-native TEZip hot path (libtezip_b200.so).
 *
 * The reference (kento/TEZip) is pure Python and has no FFI of its own; these entry points are what a
 * ctypes binding inside the reference's compress.py / decompress.py would call instead of Keras / NumPy /
 * CuPy (see INTEGRATION.md for the stubs).  Every entry point cites the reference lines it replaces
 * (paths under /root/reference/src).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - all data pointers are DEVICE pointers on the handle's / current device unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous on it
 *     unless stated otherwise.
 *   - the caller owns every buffer; the library allocates device memory only inside tz_prednet_create().
 *   - every function returns 0 on success or a negative TZ_E* code; tz_last_error() gives a thread-local
 *     message.  There is no CPU fallback: without a usable sm_100 device the calls fail with TZ_ECUDA.
 *   - frames are [n, H, W, C] u8, C fastest (the reference's channels_last layout, compress.py:116-121);
 *     predictions are [n, Hp, Wp, C] f32 with Hp, Wp = H, W rounded up to a multiple of 8
 *     (data_utils.py:77-107).
 */
#ifndef TEZIP_B200_H
#define TEZIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TZ_ABI_VERSION 2   /* 2: prev_x travels as a device pointer; 16-bit (container v2) entry points */

#define TZ_OK 0
#define TZ_EINVAL (-1)  /* bad argument */
#define TZ_ECUDA (-2)   /* CUDA runtime / driver error, or no sm_100 device */
#define TZ_ERANGE (-3)  /* data outside the domain the reference itself accepts */
#define TZ_ENOMEM (-4)

#define TZ_MAX_LAYERS 8
#define TZ_SYMBOL_OFFSET 1600 /* compress.py:348, decompress.py:236 */
#define TZ_HIST_BINS 4096     /* symbols 1600 - y must fall in [0, 4096); valid streams use 1090..2110 */

/* error-bound modes, compress.py:28-45 */
#define TZ_MODE_ABS 0
#define TZ_MODE_REL 1
#define TZ_MODE_ABSREL 2
#define TZ_MODE_PWREL 3

int tz_abi_version(void);
const char *tz_last_error(void);
/* number of usable devices with compute capability 10.x; negative on error */
int tz_device_count(void);
/* kernels launched by this library since load (all handles, this process) -- bench.py's gpu_launches */
long long tz_launch_count(void);

/* ------------------------------------------------------------------------------------------------ PredNet
 * Replaces PredNet(...) + keras Model.predict (prednet.py:24-325; call sites compress.py:195,227 and
 * decompress.py:143,154,165,175).  Inference subset only: output_mode='prediction', channels_last,
 * extrap_start_time=None, 3x3 filters, relu/relu/tanh/hard_sigmoid.
 */
typedef struct tz_prednet tz_prednet;

typedef struct tz_prednet_config {
  int n_layers;                     /* prednet.py:84 nb_layers (2..TZ_MAX_LAYERS) */
  int stack_sizes[TZ_MAX_LAYERS];   /* prednet.py:83 */
  int r_stack_sizes[TZ_MAX_LAYERS]; /* prednet.py:86 */
  int Hp, Wp;                       /* padded frame size, multiples of 2^(n_layers-1) */
  float pixel_max;                  /* prednet.py:94 */
  int max_batch;                    /* largest B ever passed to tz_prednet_next */
  int device;                       /* CUDA ordinal */
  int flags;                        /* TZ_PREDNET_* */
} tz_prednet_config;

#define TZ_PREDNET_FP32_DIRECT 1 /* fp32 CUDA-core kernels for every conv (validation path; slow) */

/* weights_host: n_weights host pointers in the reference's weight-list order (prednet.py:210-227: keys
 * a, ahat, c, f, i, o; layers ascending; kernel [3,3,Cin,Cout] then bias [Cout]); weight_elems[i] is the
 * element count of array i (checked).  Synchronous.  Pre-computes everything that does not depend on the
 * input frame (the t=0 R/C/Ahat maps, prednet.py:143-190,249-271 with zero state). */
int tz_prednet_create(const tz_prednet_config *cfg, const float *const *weights_host,
                      const long long *weight_elems, int n_weights, tz_prednet **out);
int tz_prednet_destroy(tz_prednet *h);

/* P0 = Model.predict(anything)[0,0] (compress.py:197, decompress.py:143): out f32 [Hp,Wp,C]. */
int tz_prednet_p0(tz_prednet *h, float *out, void *stream);

/* out[b] = Model.predict([in[b], zeros])[0,1] (compress.py:224-229, decompress.py:161-179) for b < B.
 * in/out f32 [B,Hp,Wp,C], contiguous; out may not alias in.  Bitwise independent of B and of which other
 * frames share the batch. */
int tz_prednet_next(tz_prednet *h, const float *in, float *out, int B, void *stream);

/* The loop of compress.py:222-229 / decompress.py:161-179 feeds every prediction straight back as the next input
 * (X_test_one = X_hat[0, 1] until the window closes).  tz_prednet_next_chained() is tz_prednet_next() whose input
 * is the first B frames of the prediction written by the previous next / next_chained call on this handle (B may
 * shrink as windows end, never grow); that buffer must still hold the prediction.  Same result bit for bit; the
 * tensor-core path has already staged the layer-0 error units of that prediction and skips one kernel and one
 * read of the frames.  Errors: no previous call, B larger than the previous B, out == the previous out. */
int tz_prednet_next_chained(tz_prednet *h, float *out, int B, void *stream);

/* Per-kernel view of one tz_prednet_next() for the roofline report: kernel i of tz_prednet_kernel_count()
 * has a name and an algorithmic FLOP count per frame; tz_prednet_next_timed() runs one next() with CUDA events
 * between the launches on `stream` and returns the device time of each kernel in ms (synchronous). */
int tz_prednet_kernel_count(tz_prednet *h);
int tz_prednet_kernel_info(tz_prednet *h, int i, char *name, int name_len, double *flops_per_frame);
int tz_prednet_next_timed(tz_prednet *h, const float *in, float *out, int B, void *stream, float *ms, int n_ms);

/* bytes of device memory the handle holds */
long long tz_prednet_device_bytes(tz_prednet *h);
/* algorithmic FLOPs per predicted frame */
double tz_prednet_flops_per_frame(tz_prednet *h);

/* ------------------------------------------------------------------------------------------------ codec ops */

/* Key-frame normalisation + padding: out[b] = pad8(lut[frames[frame_idx[b]]]) with lut[k] = f32(k)/255
 * (compress.py:138,176,219; decompress.py:117,120).  frame_idx (device, int32[B]) may be NULL = 0..B-1.
 * lut: device f32[256]. */
int tz_pad_normalize(const uint8_t *frames, const int32_t *frame_idx, const float *lut, float *out, int B,
                     int H, int W, int C, int Hp, int Wp, void *stream);

/* Residual (compress.py:293-314): x[f] = trunc_f32(pred_pool[pred_slot[f]] * 255) - frames[f] cropped to
 * HxW, or 0 where pred_slot[f] < 0 (first frame of every window, compress.py:314).
 * pred_pool f32 [slots,Hp,Wp,C]; pred_slot device int32[nt]; x int16 [nt,H,W,C]. */
int tz_residual(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, int16_t *x,
                long long nt, int H, int W, int C, int Hp, int Wp, void *stream);

/* error_bound (compress.py:23-70, driven at :315-319): in place on x for every frame f with apply[f] != 0,
 * one greedy scan per (frame, channel) plane in IEEE double, midpoint truncated toward zero.
 * mode TZ_MODE_*; b0,b1 = BOUND_VALUE[0], [1].  b0 == 0 is the identity (compress.py:24). */
int tz_error_bound(const uint8_t *frames, int16_t *x, const uint8_t *apply, long long nt, int H, int W,
                   int C, int mode, double b0, double b1, void *stream);

/* finding_difference (compress.py:73-77) fused with the symbol histogram (compress.py:348-355):
 * y[i] = x[i-1] - x[i]; y[0] = x[0] if has_prev == 0; *prev_x - x[0] if has_prev == 1 (prev_x: DEVICE pointer to the
 * last x of the previous shard -- it stays on the device, e.g. element rank-1 of an all-gathered int32 array, so that
 * no host synchronisation sits between the shards' collectives and this kernel); x[-1] - x[0] if has_prev == 2 (x
 * points into a longer device stream: chunked calls).  hist[s] += 1 for s = 1600 - y.  hist: device
 * u64[TZ_HIST_BINS], caller-zeroed.  overflow: device u64[1], caller-zeroed, counts symbols outside [0, TZ_HIST_BINS). */
int tz_delta_hist(const int16_t *x, long long n, int has_prev, const int32_t *prev_x, unsigned long long *hist,
                  unsigned long long *overflow, void *stream);

/* The last residual of a shard, x[nt*H*W*C - 1] of tz_residual() BEFORE any error bound (what a lossless shard hands
 * to its successor, compress.py:75), written to device int32 out[0]. */
int tz_last_residual(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, long long nt, int H,
                     int W, int C, int Hp, int Wp, int32_t *out, void *stream);

/* finding_difference + replacing_based_on_frequency (compress.py:84-90,339-340,348,369):
 * out[i] = lut[1600 - y[i]] (lut: device int16[TZ_HIST_BINS], symbol -> rank), or y[i] when lut == NULL
 * (the -n / ENTROPY_RUN False stream). */
/* compress.py:352-361 on the device: table[0..n) = symbols with count > 0 sorted by count descending, ties by
 * ascending symbol; lut = symbol -> rank over the 4096-symbol domain (compress.py:84-90); meta[0] = n, meta[1] = 1 if
 * some symbol lies inside [0, n) -- the reference's sequential replacement then chains and the caller must build the
 * LUT on the host (never the case for real residual streams: symbols cluster around 1600).  hist: the u64[4096]
 * histogram of tz_delta_hist / tz_encode_lossless pass 0; table, lut: int16[4096]; meta: int32[2].  Lets the rank-map
 * pass follow the histogram pass without a host round trip. */
int tz_build_table(const unsigned long long *hist, int16_t *table, int16_t *lut, int32_t *meta, void *stream);

int tz_delta_rank(const int16_t *x, long long n, int has_prev, const int32_t *prev_x, const int16_t *lut,
                  int16_t *out, void *stream);

/* Fused lossless encode (compress.py:293-314,339-369 with b0 == 0): the same results as
 * tz_residual + tz_delta_hist (pass 0) or tz_residual + tz_delta_rank (pass 1) without materialising x. */
int tz_encode_lossless(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, long long nt,
                       int H, int W, int C, int Hp, int Wp, int has_prev, const int32_t *prev_x, int pass,
                       unsigned long long *hist, unsigned long long *overflow, const int16_t *lut,
                       int16_t *out, void *stream);

/* Decoder (decompress.py:31-36,229,236 + :22-29,240-245 + :252-256,269):
 *   s = table_len >= 0 ? rank_lut[body] : body;  y = table_len >= 0 ? 1600 - s : body
 *   x[0] = first_mode == 0 ? y[0] : first_x;  x[i] = x[i-1] - y[i]     (int16 wrap-around arithmetic)
 *   out = clamp(P - x, 0, 255),  P = pred_slot[f] >= 0 ? trunc_f32(pred_pool[slot]*255) : key_plane[f]
 * rank_lut: device int16[TZ_HIST_BINS], rank -> symbol; the host builds it from the table with the
 * reference's sequential where() semantics (identity beyond the table); body values outside
 * [0, TZ_HIST_BINS) are left unchanged, as where() leaves them.  Ignored when table_len < 0.
 * workspace: device, tz_reconstruct_workspace_bytes(n) bytes.  out u8 [nt,H,W,C]; x_out optional int16[n]. */
long long tz_reconstruct_workspace_bytes(long long n);
int tz_reconstruct(const int16_t *body, long long nt, int H, int W, int C, int Hp, int Wp, int table_len,
                   const int16_t *rank_lut, int first_mode, int first_x, const float *pred_pool,
                   const int32_t *pred_slot, const uint8_t *key_plane, uint8_t *out, int16_t *x_out,
                   void *workspace, void *stream);

/* Fused lossy encode, pass 0 (compress.py:293-319 + :339-340,348-355 in ONE data pass): x = error_bound(residual)
 * is written once (int16 [nt,H,W,C]) and the histogram of the delta symbols is accumulated on the way; the rank map
 * (tz_delta_rank on x) is the second and last pass.  Same results, bit for bit, as tz_residual + tz_error_bound +
 * tz_delta_hist.  apply: device u8[nt] as in tz_error_bound.  counter: device u32[1], caller-zeroed (the last CTA to
 * finish adds the symbols that straddle frame boundaries).  has_prev: 0, 1 (*prev_x, device) as in tz_delta_hist, or
 * 3 = the caller accounts for the first symbol of the stream itself (a shard whose halo arrives later: follow up with
 * tz_delta_hist(x, 1, has_prev, prev_x, ...)).  tz_encode_lossy_supported(): plane-wide bounds (abs / rel / absrel),
 * 1..4 channels, W*C a multiple of 8; otherwise use the separate entry points. */
int tz_encode_lossy_supported(int H, int W, int C, int mode);
int tz_encode_lossy(const uint8_t *frames, const float *pred_pool, const int32_t *pred_slot, const uint8_t *apply,
                    int16_t *x, long long nt, int H, int W, int C, int Hp, int Wp, int mode, double b0, double b1,
                    int has_prev, const int32_t *prev_x, unsigned long long *hist, unsigned long long *overflow,
                    unsigned int *counter, void *stream);

/* ------------------------------------------------------------------------------------------------ 16-bit samples
 * Container v2 (BASELINE config 4: 1024x1024x1 u16 frames).  The reference cannot hold
 * such data: compress.py:106-110 turns every input into 8-bit RGB, :183 keeps a u8 key plane, :333 int16 residuals,
 * :348 the symbol offset 1600, :394 int16 shape fields.  These entry points restate the same pipeline at the widths
 * 16-bit samples need -- the 8-bit container and its entry points above are untouched:
 *   x = trunc_f32(pred * 65535) - sample (int32)      y = delta as compress.py:73-77 (int32)
 *   s = TZ_WIDE_OFFSET - y, histogram bin = s - TZ_WIDE_SYM_MIN in [0, TZ_WIDE_BINS)
 *   table / ranks as compress.py:352-369, int32.  TZ_WIDE_OFFSET follows the reference's rule for 1600: every
 *   symbol (>= 268930) lies above every possible rank (<= 262140), so the replacement is a pure look-up.
 * frames / key plane: u16 [n,H,W,C]; predictions as above.  prev_x / has_prev as in tz_delta_hist. */
#define TZ_WIDE_OFFSET 400000
#define TZ_WIDE_SYM_MIN (TZ_WIDE_OFFSET - 131071)
#define TZ_WIDE_BINS 262144

/* out[b] = pad8(f32(frames[frame_idx[b]]) / 65535) (compress.py:138,176 with the 16-bit maximum; IEEE division) */
int tz_pad_normalize16(const uint16_t *frames, const int32_t *frame_idx, float *out, int B, int H, int W, int C,
                       int Hp, int Wp, void *stream);
/* compress.py:293-314 -> x int32 [nt,H,W,C] (needed only for the lossy modes; lossless encodes fused) */
int tz_residual16(const uint16_t *frames, const float *pred_pool, const int32_t *pred_slot, int32_t *x, long long nt,
                  int H, int W, int C, int Hp, int Wp, void *stream);
int tz_last_residual16(const uint16_t *frames, const float *pred_pool, const int32_t *pred_slot, long long nt, int H,
                       int W, int C, int Hp, int Wp, int32_t *out, void *stream);
/* compress.py:23-70 in place on int32 residual planes, bounds in 0..65535 level units */
int tz_error_bound16(const uint16_t *frames, int32_t *x, const uint8_t *apply, long long nt, int H, int W, int C,
                     int mode, double b0, double b1, void *stream);
/* compress.py:339-369.  pass 0: histogram of the delta symbols into hist (device u64[TZ_WIDE_BINS], caller-zeroed;
 * overflow u64[1]); pass 1: out[i] = lut[bin(y[i])] (lut: device int32[TZ_WIDE_BINS] from tz_build_table16), or y[i]
 * when lut == NULL (-n streams).  x != NULL: the materialised (error-bounded) residual is the source; x == NULL: the
 * residual is recomputed from frames + pred_pool + pred_slot (lossless: nothing but the codes is ever written). */
int tz_encode16(const uint16_t *frames, const float *pred_pool, const int32_t *pred_slot, const int32_t *x,
                long long nt, int H, int W, int C, int Hp, int Wp, int has_prev, const int32_t *prev_x, int pass,
                unsigned long long *hist, unsigned long long *overflow, const int32_t *lut, int32_t *out,
                void *stream);
/* compress.py:352-361 on the device: table[0..n) = symbols by count descending, ties ascending; lut = bin -> rank
 * (identity symbol elsewhere); meta[0] = n (meta: device int32[2]).  workspace: tz_build_table16_workspace_bytes(). */
long long tz_build_table16_workspace_bytes(void);
int tz_build_table16(const unsigned long long *hist, int32_t *table, int32_t *lut, int32_t *meta, void *workspace,
                     void *stream);
/* decoder, as tz_reconstruct: rank_lut device int32[TZ_WIDE_BINS] (rank -> symbol, identity beyond the table),
 * y = TZ_WIDE_OFFSET - s, out = clamp(P - x, 0, 65535) with P = trunc_f32(pred * 65535) or the key sample */
long long tz_reconstruct16_workspace_bytes(long long n);
int tz_reconstruct16(const int32_t *body, long long nt, int H, int W, int C, int Hp, int Wp, int table_len,
                     const int32_t *rank_lut, int first_mode, int first_x, const float *pred_pool,
                     const int32_t *pred_slot, const uint16_t *key_plane, uint16_t *out, int32_t *x_out,
                     void *workspace, void *stream);
/* compress.py:245-246 on 16-bit samples (f32(sample) / 65535 against the prediction, float64 sums) */
int tz_window_sse16(const uint16_t *frames, const int32_t *frame_idx, const float *pred, double *sse, int B, int H,
                    int W, int C, int Hp, int Wp, void *stream);

/* DWP metric (compress.py:245-246): sse[b] = sum over the padded Hp x Wp x C area of
 * (f64(lut[frame]) - f64(pred))^2 with zeros in the padding, fixed reduction order (deterministic).
 * frames u8 [*,H,W,C]; frame_idx device int32[B]; pred f32 [B,Hp,Wp,C]; sse device f64[B]. */
int tz_window_sse(const uint8_t *frames, const int32_t *frame_idx, const float *lut, const float *pred,
                  double *sse, int B, int H, int W, int C, int Hp, int Wp, void *stream);

/* Dynamic windows without a host round trip per step (compress.py:214-266 for B chains in lock step; the state lives
 * in device arrays of B entries).  tz_dwp_gather builds the inputs of the next PredNet step: X[b] =
 * pad8(f32(frames[key[b]]) / pixel_max) if idx[b] == key[b] + 1 (compress.py:219), else pred_pool[last[b]] (:222).
 * frames: u8 (bits = 8) or u16 (bits = 16).  tz_dwp_update takes the close decision of :245-263 per chain after
 * tz_window_sse (sse_step[b] of frame idx[b]): cumulative mean squared error over the window > threshold (or the
 * static (idx - p) % window == 0 test) -> idx[b] becomes a key (is_key[idx] = 1, pred_slot[idx] = -1), else
 * pred_slot[idx] = slot0 + b, apply[idx] = 1; idx[b] += 1.  denom = Hp*Wp*C. */
int tz_dwp_gather(const void *frames, int bits, const float *pred_pool, const int32_t *key, const int32_t *idx,
                  const int32_t *last, float *X, int B, int H, int W, int C, int Hp, int Wp, void *stream);
int tz_dwp_update(const double *sse_step, int32_t *key, int32_t *idx, int32_t *last, double *sse, int32_t *cnt,
                  int32_t *pred_slot, uint8_t *apply, uint8_t *is_key, int B, int slot0, double denom,
                  int has_threshold, double threshold, int window, int p, void *stream);

/* key-frame plane (compress.py:183,190,220,261): out[f] = is_key[f] ? frames[f] : 0. */
int tz_key_plane(const uint8_t *frames, const uint8_t *is_key, uint8_t *out, long long nt,
                 long long frame_bytes, void *stream);

/* decompress.py:123-127: nonzero[f] = any(key_plane[f] != 0). nonzero: device u8[nt]. */
int tz_frames_nonzero(const uint8_t *key_plane, uint8_t *nonzero, long long nt, long long frame_bytes,
                      void *stream);

/* Strided frame copy between host (pinned) and device, either direction: `height` pieces of `width` bytes, piece k
 * at dst + k*dpitch from src + k*spitch.  Lets a caller send the key frames of every window first (the prediction
 * steps need nothing else -- compress.py:219-229 reads one frame per window) and the remaining frames behind the
 * PredNet kernels, where the reference loads the whole array up front (compress.py:131-152).  One 2-D DMA
 * (cudaMemcpy2DAsync, direction inferred from the pointers); asynchronous with respect to the host when the host
 * side is pinned. */
int tz_memcpy2d_async(void *dst, long long dpitch, const void *src, long long spitch, long long width,
                      long long height, void *stream);

/* ---- the container's lossless back-end: compress.py:276 (`zstd.compress(key_frame_str, 9)`) and compress.py:398
 * (`zstd.compress(entropy bytes, 9)`) -- the reference's decoder (decompress.py:89,98 `zstd.decompress`) needs one zstd
 * frame with its content size, nothing else.  These two calls write such a frame from `n` bytes that are already in
 * device memory: per 128 KB block an RLE block (one repeated byte), a raw block, or a compressed block that holds the
 * block Huffman-coded as four literal streams and no sequences (RFC 8878 3.1.1.3).  The Huffman code is the caller's
 * (host side: tezip_b200/zstd_frames.py builds it from the histogram of tz_zstd_hist).
 *
 * tz_zstd_hist:   hist: device u32[256], byte counts of the blocks that are neither uniform nor shorter than 1024;
 *                 uniform: device i32[ceil(n / 131072)], the byte value of a block that holds one value only, else -1.
 * tz_zstd_encode: ctable: device u32[256], code | nbits << 16 (nbits <= 11; canonical order of the zstd decoder);
 *                 tree: device bytes of the Huffman_Tree_Description, tree_len of them (0: no code, blocks are RLE or
 *                 raw); uniform: as written by tz_zstd_hist; workspace: tz_zstd_workspace_bytes(n) device bytes;
 *                 out: tz_zstd_bound(n) device bytes, 4-byte aligned; total: device u64, the frame's size in bytes. */
unsigned long long tz_zstd_bound(unsigned long long n);
unsigned long long tz_zstd_workspace_bytes(unsigned long long n);
int tz_zstd_hist(const uint8_t *src, unsigned long long n, uint32_t *hist, int32_t *uniform, void *stream);
int tz_zstd_encode(const uint8_t *src, unsigned long long n, const uint32_t *ctable, const uint8_t *tree,
                   unsigned tree_len, const int32_t *uniform, void *workspace, uint8_t *out,
                   unsigned long long *total, void *stream);

/* decompress.py:89,98 (`zstd.decompress`) for frames of that subset: the host walks the frame and block headers
 * (tezip_b200/zstd_frames.py parse_frame; any other frame is decoded by libzstd as in the reference) and the device
 * decodes all blocks, and the four Huffman streams of each, at once.
 * frame: the compressed frame in device memory; blocks: device array of nblocks records {u64 src_off, u64 dst_off,
 * u32 type (0 raw, 1 RLE, 2 Huffman), u32 regen, u32 stream_bytes[4], u32 table, u32 pad} (48 bytes each); dtables:
 * device u16[T][2048], symbol | nbits << 8 indexed by the next 11 bits (NULL if no block is Huffman-coded); out: the
 * decoded content; err: device i32, 0 or the code of a malformed stream. */
int tz_zstd_decode(const uint8_t *frame, const void *blocks, unsigned long long nblocks, const uint16_t *dtables,
                   uint8_t *out, int32_t *err, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TEZIP_B200_H */
