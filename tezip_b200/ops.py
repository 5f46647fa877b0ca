"""Device-side codec ops: thin torch-tensor wrappers over the C ABI (include/tezip_b200.h).

Names follow the reference's module-level functions (compress.py:23-90, decompress.py:22-36) so parity tests
read like the reference's code.  All tensors are CUDA tensors; every op runs on torch's current stream.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import MODES, TZ_HIST_BINS, TZ_SYMBOL_OFFSET, TZ_WIDE_BINS, TZ_WIDE_OFFSET, TZ_WIDE_SYM_MIN, check, ptr


def _st(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def is_wide(t):
    """True for the 16-bit sample / int32 code family (container v2), False for the reference's u8 / int16."""
    return t.dtype in (torch.uint16, torch.int32)


def _prev(dev, has_prev, prev_x):
    """(has_prev, device pointer) for the delta halo: prev_x may be a device int32 tensor (stays on the device, the
    sharded path) or a Python int (tests, single calls)."""
    hp = int(has_prev)                      # 0 / 1 / 2 (chunked) / 3 (tz_encode_lossy: first symbol by the caller)
    if hp != 1:
        return hp, None, None
    if not torch.is_tensor(prev_x):
        prev_x = torch.tensor([int(prev_x)], dtype=torch.int32, device=dev)
    assert prev_x.is_cuda and prev_x.dtype == torch.int32 and prev_x.numel() >= 1
    return hp, ptr(prev_x), prev_x     # the tensor is returned so that the caller keeps it alive across the launch


_LUT_CACHE = {}


def norm_lut(device):
    """lut[k] = float32(k)/255 as compress.py:138 computes it (identical to decompress.py:117)."""
    key = str(device)
    if key not in _LUT_CACHE:
        _LUT_CACHE[key] = torch.from_numpy(np.arange(256).astype(np.float32) / 255).to(device)
    return _LUT_CACHE[key]


def pad_normalize(frames, frame_idx, Hp, Wp, out=None):
    """frames u8 [n,H,W,C]; frame_idx int32 [B] (device) or None -> f32 [B,Hp,Wp,C]."""
    n, H, W, C = frames.shape
    B = n if frame_idx is None else frame_idx.numel()
    if out is None:
        out = torch.empty((B, Hp, Wp, C), dtype=torch.float32, device=frames.device)
    if is_wide(frames):
        check(_lib.load().tz_pad_normalize16(ptr(frames), ptr(frame_idx), ptr(out), B, H, W, C, Hp, Wp,
                                             _st(frames.device)), "tz_pad_normalize16")
        return out
    check(_lib.load().tz_pad_normalize(ptr(frames), ptr(frame_idx), ptr(norm_lut(frames.device)), ptr(out), B, H, W, C,
                                       Hp, Wp, _st(frames.device)), "tz_pad_normalize")
    return out


def residual(frames, pred_pool, pred_slot, out=None):
    """compress.py:293-314 -> int16 [n,H,W,C] (int32 for 16-bit samples)."""
    n, H, W, C = frames.shape
    _s, Hp, Wp, _c = pred_pool.shape
    if is_wide(frames):
        if out is None:
            out = torch.empty((n, H, W, C), dtype=torch.int32, device=frames.device)
        check(_lib.load().tz_residual16(ptr(frames), ptr(pred_pool), ptr(pred_slot), ptr(out), n, H, W, C, Hp, Wp,
                                        _st(frames.device)), "tz_residual16")
        return out
    if out is None:
        out = torch.empty((n, H, W, C), dtype=torch.int16, device=frames.device)
    check(_lib.load().tz_residual(ptr(frames), ptr(pred_pool), ptr(pred_slot), ptr(out), n, H, W, C, Hp, Wp,
                                  _st(frames.device)), "tz_residual")
    return out


def error_bound(frames, x, apply, mode, value):
    """compress.py:23-70 applied in place to every flagged frame of x (int16 [n,H,W,C])."""
    n, H, W, C = frames.shape
    b0 = float(value[0])
    b1 = float(value[1]) if len(value) > 1 else 0.0
    if is_wide(frames):
        check(_lib.load().tz_error_bound16(ptr(frames), ptr(x), ptr(apply), n, H, W, C, MODES[mode], b0, b1,
                                           _st(frames.device)), "tz_error_bound16")
        return x
    check(_lib.load().tz_error_bound(ptr(frames), ptr(x), ptr(apply), n, H, W, C, MODES[mode], b0, b1,
                                     _st(frames.device)), "tz_error_bound")
    return x


def finding_difference_hist(x, hist, overflow, has_prev=False, prev_x=0):
    """compress.py:73-77 + :348-355: accumulates the histogram of 1600 - y into hist (u64 as int64[4096])."""
    hp, pp, _keep = _prev(x.device, has_prev, prev_x)
    check(_lib.load().tz_delta_hist(ptr(x), x.numel(), hp, pp, ptr(hist), ptr(overflow), _st(x.device)),
          "tz_delta_hist")


def finding_difference_rank(x, lut, out=None, has_prev=False, prev_x=0):
    """compress.py:73-77 + replacing_based_on_frequency (:84-90); lut None -> the raw delta stream."""
    if out is None:
        out = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
    if out is not None and out.numel() != x.numel():
        raise ValueError("out must have as many elements as x")
    hp, pp, _keep = _prev(x.device, has_prev, prev_x)
    check(_lib.load().tz_delta_rank(ptr(x), x.numel(), hp, pp, ptr(lut), ptr(out), _st(x.device)), "tz_delta_rank")
    return out


def encode_lossless(frames, pred_pool, pred_slot, pass_, hist=None, overflow=None, lut=None, out=None,
                    has_prev=False, prev_x=0):
    n, H, W, C = frames.shape
    _s, Hp, Wp, _c = pred_pool.shape
    hp, pp, _keep = _prev(frames.device, has_prev, prev_x)
    check(_lib.load().tz_encode_lossless(ptr(frames), ptr(pred_pool), ptr(pred_slot), n, H, W, C, Hp, Wp, hp, pp,
                                         pass_, ptr(hist), ptr(overflow), ptr(lut), ptr(out), _st(frames.device)),
          "tz_encode_lossless")
    return out


def encode_lossy_supported(frames, mode):
    _n, H, W, C = frames.shape
    return (not is_wide(frames)) and bool(_lib.load().tz_encode_lossy_supported(H, W, C, MODES[mode]))


def encode_lossy(frames, pred_pool, pred_slot, apply, mode, value, hist, overflow, counter, x=None, has_prev=False,
                 prev_x=0):
    """compress.py:293-319 + :339-340,348-355 in one data pass (tz_encode_lossy): returns x = error_bound(residual)
    (int16 [n,H,W,C]) and accumulates the delta-symbol histogram.  counter: device int32[1], zeroed by the caller."""
    n, H, W, C = frames.shape
    _s, Hp, Wp, _c = pred_pool.shape
    if x is None:
        x = torch.empty((n, H, W, C), dtype=torch.int16, device=frames.device)
    b0 = float(value[0])
    b1 = float(value[1]) if len(value) > 1 else 0.0
    hp, pp, _keep = _prev(frames.device, has_prev, prev_x)
    check(_lib.load().tz_encode_lossy(ptr(frames), ptr(pred_pool), ptr(pred_slot), ptr(apply), ptr(x), n, H, W, C, Hp,
                                      Wp, MODES[mode], b0, b1, hp, pp, ptr(hist), ptr(overflow), ptr(counter),
                                      _st(frames.device)), "tz_encode_lossy")
    return x


def encode16(frames, pred_pool, pred_slot, x, pass_, hist=None, overflow=None, lut=None, out=None, has_prev=False,
             prev_x=0):
    """tz_encode16 (16-bit samples): pass 0 = histogram, pass 1 = rank map (or the raw delta stream, lut None); the
    source is the materialised residual x (int32) when given, else frames + predictions (fused lossless)."""
    n, H, W, C = frames.shape
    _s, Hp, Wp, _c = pred_pool.shape
    hp, pp, _keep = _prev(frames.device, has_prev, prev_x)
    check(_lib.load().tz_encode16(ptr(frames), ptr(pred_pool), ptr(pred_slot), ptr(x), n, H, W, C, Hp, Wp, hp, pp,
                                  pass_, ptr(hist), ptr(overflow), ptr(lut), ptr(out), _st(frames.device)),
          "tz_encode16")
    return out


def finding_difference_rank16(x, lut, out, has_prev=False, prev_x=0):
    """tz_encode16 pass 1 on a flat int32 residual stream (or a chunk of one: has_prev == 2)."""
    hp, pp, _keep = _prev(x.device, has_prev, prev_x)
    n = x.numel()
    check(_lib.load().tz_encode16(None, None, None, ptr(x), n, 1, 1, 1, 1, 1, hp, pp, 1, None, None, ptr(lut), ptr(out),
                                  _st(x.device)), "tz_encode16")
    return out


def last_residual(frames, pred_pool, pred_slot, out=None):
    """The last residual of a shard as a device int32[1] (the delta halo its successor needs, compress.py:75)."""
    n, H, W, C = frames.shape
    _s, Hp, Wp, _c = pred_pool.shape
    if out is None:
        out = torch.empty(1, dtype=torch.int32, device=frames.device)
    fn = "tz_last_residual16" if is_wide(frames) else "tz_last_residual"
    check(getattr(_lib.load(), fn)(ptr(frames), ptr(pred_pool), ptr(pred_slot), n, H, W, C, Hp, Wp, ptr(out),
                                   _st(frames.device)), fn)
    return out


_TABLE16_WS = {}


def build_table16_device(hist, table, lut, meta):
    """compress.py:352-361 for the wide symbol range (tz_build_table16): hist int64[TZ_WIDE_BINS] -> table int32,
    lut int32[TZ_WIDE_BINS] (bin -> rank), meta int32[2] = (table length, 0)."""
    lib = _lib.load()
    key = str(hist.device)
    if key not in _TABLE16_WS:
        _TABLE16_WS[key] = torch.empty(int(lib.tz_build_table16_workspace_bytes()), dtype=torch.uint8, device=hist.device)
    check(lib.tz_build_table16(ptr(hist), ptr(table), ptr(lut), ptr(meta), ptr(_TABLE16_WS[key]), _st(hist.device)),
          "tz_build_table16")


def build_table16(hist_np):
    """Host form of the wide table: symbols (TZ_WIDE_SYM_MIN + bin) by count descending, ties ascending."""
    ii = np.nonzero(hist_np)[0]
    order = np.lexsort((ii, -hist_np[ii].astype(np.int64)))
    return (ii[order] + TZ_WIDE_SYM_MIN).astype(np.int32)


def decode_lut16(table):
    """rank -> symbol over [0, TZ_WIDE_BINS), identity beyond the table (decompress.py:31-36; the wide offset keeps
    every symbol above every rank, so the sequential replacement is a plain scatter)."""
    t = np.asarray(table).astype(np.int64)
    if len(t) and (int(t.min()) < len(t) or len(np.unique(t)) != len(t)):
        raise _lib.TezipError("corrupt wide table: a symbol lies inside the rank range")
    lut = np.arange(TZ_WIDE_BINS, dtype=np.int32)
    lut[:len(t)] = t.astype(np.int32)
    return lut


def build_table_device(hist, table, lut, meta):
    """compress.py:352-361 + :84-90 on the device (tz_build_table): hist int64[4096] -> table int16[4096],
    lut int16[4096], meta int32[2] = (table length, 1 if the LUT must be rebuilt on the host)."""
    check(_lib.load().tz_build_table(ptr(hist), ptr(table), ptr(lut), ptr(meta), _st(hist.device)), "tz_build_table")


def build_table(hist_np):
    """compress.py:352-361: symbols with count > 0 sorted by count descending, ties by ascending symbol."""
    ii = np.nonzero(hist_np)[0]
    order = np.lexsort((ii, -hist_np[ii].astype(np.int64)))
    return ii[order].astype(np.int16)


def _sequential_replace(pairs):
    """lut after `for (src, dst) in pairs: lut[lut == src] = dst`, starting from the identity over the 4096-symbol
    domain.  The reference runs one full-array where() per table entry (compress.py:84-90, decompress.py:31-36); a
    later pass sees the values earlier passes wrote, so value/index collisions chain.  Here the positions that
    currently hold a value are kept as a group and whole groups move, which is the same function in O(len(table))
    instead of O(len(table) * 4096) -- the table is on the critical path between the two GPU passes."""
    groups = {}

    def group(v):
        g = groups.get(v)
        if g is None:
            g = groups[v] = [v] if 0 <= v < TZ_HIST_BINS else []
        return g

    for src, dst in pairs:
        if src == dst:
            continue
        g = group(src)
        if g:
            group(dst).extend(g)
            groups[src] = []
    pos, val = [], []
    for v, g in groups.items():
        pos.extend(g)
        val.extend([v] * len(g))
    lut = np.arange(TZ_HIST_BINS, dtype=np.int16)
    if pos:
        lut[pos] = np.array(val, dtype=np.int64).astype(np.int16)
    return lut


def _no_collisions(t):
    """True when every symbol of the table lies outside the rank range [0, len(table)) (and inside the domain) and
    no symbol repeats: then no where() pass can see a value written by another pass, and the sequential replace
    is a plain scatter.  The usual case: symbols cluster around 1600, tables hold tens of entries."""
    return len(t) > 0 and int(t.min()) >= len(t) and int(t.max()) < TZ_HIST_BINS and len(np.unique(t)) == len(t)


def encode_lut(table):
    """symbol -> rank over the whole 4096-symbol domain, equal to the reference's sequential where() passes
    (compress.py:84-90: result[result == num] = idx for every table entry in order)."""
    t = np.asarray(table).astype(np.int64)
    if _no_collisions(t):
        lut = np.arange(TZ_HIST_BINS, dtype=np.int16)
        lut[t] = np.arange(len(t), dtype=np.int16)
        return lut
    return _sequential_replace((num, idx) for idx, num in enumerate(t.tolist()))


def decode_lut(table):
    """rank -> symbol (decompress.py:31-36: result[result == idx] = num in table order), identity beyond the table."""
    t = np.asarray(table).astype(np.int64)
    if _no_collisions(t):
        lut = np.arange(TZ_HIST_BINS, dtype=np.int16)
        lut[:len(t)] = t.astype(np.int16)
        return lut
    return _sequential_replace((idx, num) for idx, num in enumerate(t.tolist()))


def reconstruct(body, shape, Hp, Wp, table_len, rank_lut, pred_pool, pred_slot, key_plane, first_mode=0, first_x=0,
                want_x=False, out=None):
    """decompress.py:229,236,240-245,252-256,269 -> u8 [n,H,W,C] (and x int16 if want_x).
    out: optional preallocated result (same shape and sample type), e.g. a staging buffer the caller recycles."""
    n, H, W, C = shape
    dev = body.device
    lib = _lib.load()
    if out is not None:
        want = torch.uint16 if is_wide(body) else torch.uint8
        if tuple(out.shape) != (n, H, W, C) or out.dtype != want or not out.is_contiguous() or out.device != dev:
            raise ValueError("out must be a contiguous %s [%d,%d,%d,%d] tensor on %s" % (want, n, H, W, C, dev))
    if is_wide(body):
        ws = torch.empty(int(lib.tz_reconstruct16_workspace_bytes(n * H * W * C)), dtype=torch.uint8, device=dev)
        if out is None:
            out = torch.empty((n, H, W, C), dtype=torch.uint16, device=dev)
        x = torch.empty(n * H * W * C, dtype=torch.int32, device=dev) if want_x else None
        check(lib.tz_reconstruct16(ptr(body), n, H, W, C, Hp, Wp, int(table_len), ptr(rank_lut), int(first_mode),
                                   int(first_x), ptr(pred_pool), ptr(pred_slot), ptr(key_plane), ptr(out), ptr(x),
                                   ptr(ws), _st(dev)), "tz_reconstruct16")
        return (out, x) if want_x else out
    ws = torch.empty(int(lib.tz_reconstruct_workspace_bytes(n * H * W * C)), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((n, H, W, C), dtype=torch.uint8, device=dev)
    x = torch.empty(n * H * W * C, dtype=torch.int16, device=dev) if want_x else None
    check(lib.tz_reconstruct(ptr(body), n, H, W, C, Hp, Wp, int(table_len), ptr(rank_lut), int(first_mode),
                             int(first_x), ptr(pred_pool), ptr(pred_slot), ptr(key_plane), ptr(out), ptr(x), ptr(ws),
                             _st(dev)), "tz_reconstruct")
    return (out, x) if want_x else out


def window_sse(frames, frame_idx, pred, out=None):
    """compress.py:245-246 numerator per chain: f64 [B]."""
    _n, H, W, C = frames.shape
    B, Hp, Wp, _c = pred.shape
    if out is None:
        out = torch.empty(B, dtype=torch.float64, device=frames.device)
    if is_wide(frames):
        check(_lib.load().tz_window_sse16(ptr(frames), ptr(frame_idx), ptr(pred), ptr(out), B, H, W, C, Hp, Wp,
                                          _st(frames.device)), "tz_window_sse16")
        return out
    check(_lib.load().tz_window_sse(ptr(frames), ptr(frame_idx), ptr(norm_lut(frames.device)), ptr(pred), ptr(out), B,
                                    H, W, C, Hp, Wp, _st(frames.device)), "tz_window_sse")
    return out


def dwp_gather(frames, pred_pool, key, idx, last, X, B):
    """Inputs of the next DWP step for B chains (tz_dwp_gather): X[b] = the normalised, padded key frame of chain b
    if it has just opened a window (compress.py:219), else its previous prediction pool[last[b]] (compress.py:222)."""
    _n, H, W, C = frames.shape
    _s, Hp, Wp, _c = pred_pool.shape
    check(_lib.load().tz_dwp_gather(ptr(frames), 16 if is_wide(frames) else 8, ptr(pred_pool), ptr(key), ptr(idx),
                                    ptr(last), ptr(X), int(B), H, W, C, Hp, Wp, _st(frames.device)), "tz_dwp_gather")
    return X


def dwp_update(sse_step, key, idx, last, sse, cnt, pred_slot, apply, is_key, B, slot0, denom, threshold, window, p):
    """The close decision of compress.py:245-263 for B chains on the device (tz_dwp_update)."""
    check(_lib.load().tz_dwp_update(ptr(sse_step), ptr(key), ptr(idx), ptr(last), ptr(sse), ptr(cnt), ptr(pred_slot),
                                    ptr(apply), ptr(is_key), int(B), int(slot0), float(denom),
                                    0 if threshold is None else 1, 0.0 if threshold is None else float(threshold),
                                    0 if window is None else int(window), int(p), _st(sse_step.device)),
          "tz_dwp_update")


def key_plane(frames, is_key, out=None):
    n = frames.shape[0]
    fb = frames[0].numel() * frames.element_size()
    if out is None:
        out = torch.empty_like(frames)
    check(_lib.load().tz_key_plane(ptr(frames), ptr(is_key), ptr(out), n, fb, _st(frames.device)), "tz_key_plane")
    return out


def frames_nonzero(plane):
    n = plane.shape[0]
    fb = plane[0].numel() * plane.element_size()
    out = torch.empty(n, dtype=torch.uint8, device=plane.device)
    check(_lib.load().tz_frames_nonzero(ptr(plane), ptr(out), n, fb, _st(plane.device)), "tz_frames_nonzero")
    return out
