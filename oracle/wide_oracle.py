"""ORACLE (test infrastructure only) -- CPU restatement of the hot path for 16-bit samples (container v2,
BASELINE config 4).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this.

The reference CANNOT run this case: compress.py:106-110 converts every image to 8-bit RGB, :183 keeps a u8 key plane,
:333 casts residuals to int16, :348 uses the symbol offset 1600 (room for |y| <= 510 only), :394 writes int16 shape
fields.  PARITY UNPINNED against the reference for that reason.  What pins this file instead: it is the SAME code path
as oracle/codec_oracle.py -- which is pinned byte for byte against the unmodified reference -- with four constants
widened, and tests/test_wide_oracle.py checks that with 8-bit data and the reference's constants these functions
reproduce codec_oracle's (hence the reference's) stream exactly.  The widened constants:
    pixel maximum      255  -> 65535     compress.py:138,294,307-308; decompress.py:117,252,255
    residual / codes   int16 -> int32    compress.py:333,394
    symbol offset      1600 -> 400000    compress.py:348; decompress.py:236 (same rule: every symbol above every rank)
    key plane          u8   -> u16       compress.py:183
and the trailer of the stream: [codes] [table] [T | -1] [1, nt, H, W, C] [p] [bits = 16] [version = 2] [magic].
"""
import time

import numpy as np

from . import codec_oracle as co

OFFSET = 400000            # include/tezip_b200.h TZ_WIDE_OFFSET
SYM_MIN = OFFSET - 131071
NBINS = 262144
MAGIC = 0x5A543230
VERSION = 2


class Width:
    """The constants that differ between the reference's 8-bit pipeline and the 16-bit extension."""

    def __init__(self, pixmax, code, offset, pix):
        self.pixmax, self.code, self.offset, self.pix = pixmax, np.dtype(code), offset, np.dtype(pix)


W8 = Width(255, np.int16, 1600, np.uint8)          # the reference
W16 = Width(65535, np.int32, OFFSET, np.uint16)    # container v2


def delta_encode(x):
    """compress.py:73-77 (vectorised; the arithmetic wraps in the code type exactly as the reference's int16 does)."""
    x = np.ascontiguousarray(x).ravel()
    y = np.empty_like(x)
    if x.size:
        y[0] = x[0]
        y[1:] = x[:-1] - x[1:]
    return y


def delta_decode(y):
    """decompress.py:22-29: x[0] = y[0], x[i] = x[i-1] - y[i]  ==  y[0] - cumsum(y[1:]) in wrapping arithmetic."""
    y = np.ascontiguousarray(y).ravel()
    if not y.size:
        return y.copy()
    u = y.astype(np.int64)
    x = u[0] - np.concatenate([[0], np.cumsum(u[1:])])
    return x.astype(y.dtype)          # truncating cast == wrap-around


def build_table(s, width):
    """compress.py:352-361: symbols with count > 0 by count descending; the stable sort keeps ties ascending."""
    lo = int(s.min())
    cnt = np.bincount((s.astype(np.int64) - lo))
    ii = np.nonzero(cnt)[0]
    d = list(zip(ii + lo, cnt[ii]))
    d.sort(key=lambda e: e[1], reverse=True)
    return np.array([k for k, _ in d], dtype=width.code)


def replacing_encode(s, table):
    """compress.py:84-90.  The reference runs one where() pass per table entry; when no symbol lies inside the rank
    range [0, len(table)) no pass can see a value an earlier pass wrote, so the passes are one look-up
    (SURVEY.md A13).  That holds by construction for both offsets (1090 > 1020, 268930 > 262140) and is asserted."""
    t = table.astype(np.int64)
    assert len(t) == 0 or int(t.min()) >= len(t), "symbol inside the rank range: sequential semantics needed"
    order = np.argsort(t, kind="stable")
    pos = np.searchsorted(t[order], s.astype(np.int64))
    assert np.array_equal(t[order][pos], s.astype(np.int64))
    return order[pos].astype(table.dtype)


def replacing_decode(body, table):
    """decompress.py:31-36 as a look-up (same argument); values beyond the table are left alone."""
    t = table.astype(np.int64)
    assert len(t) == 0 or int(t.min()) >= len(t)
    b = body.astype(np.int64)
    inside = (b >= 0) & (b < len(t))
    out = b.copy()
    out[inside] = t[b[inside]]
    return out.astype(body.dtype)


def encode_windows(origine_img, windows, PRE, mode, bound, entropy=True, width=W16, timers=None):
    """compress.py:289-395 with the widths of `width` (see codec_oracle.encode_windows, line for line)."""
    T = timers if timers is not None else {}
    tic = time.perf_counter
    shape = origine_img.shape
    H, W = shape[2], shape[3]
    difference_list = []
    t0 = tic()
    t_eb = 0.0
    for wi, (first, preds) in enumerate(windows):
        n = len(preds)
        origine_pick = origine_img[:, first:first + n] / width.pixmax           # :294
        predict_pick = np.stack(preds)[np.newaxis][:, :, :H, :W]               # :295-298
        X_hat_1 = np.multiply(predict_pick, np.float32(width.pixmax))          # :307 float32 product
        assert X_hat_1.dtype == np.float32
        X_test_1 = np.multiply(origine_pick, float(width.pixmax)).astype(int)   # :308,310 (exact for every sample value)
        X_hat_1 = X_hat_1.astype(int)                                           # :311
        difference = X_hat_1 - X_test_1                                         # :313
        difference[:, 0] = 0                                                    # :314
        if not (PRE != 0 and wi == 0):                                          # :315
            t1 = tic()
            d0 = np.ascontiguousarray(difference[0])
            co.error_bound_frames(np.ascontiguousarray(X_test_1[0]), d0, mode, list(bound))   # :316-319
            difference = d0[np.newaxis]
            t_eb += tic() - t1
        difference_list.append(difference)
    T["residual"] = tic() - t0 - t_eb
    T["error_bound"] = t_eb
    x = np.concatenate(difference_list, axis=1).astype(width.code).ravel()     # :329-333
    t0 = tic()
    y = delta_encode(x)                                                         # :339-340
    T["finding_difference"] = tic() - t0
    result, table = y, None
    if entropy:
        t0 = tic()
        s = np.subtract(width.code.type(width.offset), y)                       # :348
        table = build_table(s, width)                                           # :352-361
        T["table_create"] = tic() - t0
        t0 = tic()
        result = replacing_encode(s, table)                                     # :369
        T["replacing"] = tic() - t0
    tail = ([int(v) for v in table] + [len(table)]) if entropy else [-1]        # :383-387
    tail += [int(v) for v in shape] + [PRE]                                     # :390-392
    if width is W16:
        tail += [16, VERSION, MAGIC]
    payload = np.concatenate([result.astype(np.int64), np.array(tail, np.int64)]).astype(width.code)   # :394
    preds_full = np.concatenate([np.stack(w[1]) for w in windows], axis=0)
    return {"payload": payload, "windows": [(w[0], len(w[1])) for w in windows], "preds": preds_full,
            "x": x, "y": y, "table": table, "shape": tuple(shape)}


def compress_arrays(frames, predictor, p, window, threshold, mode, bound, entropy=True, width=W16, timers=None):
    """compress.py:138-395 on arrays (the scheduler of codec_oracle.compress_arrays with `width`'s pixel maximum)."""
    T = timers if timers is not None else {}
    tic = time.perf_counter
    origine_img = np.ascontiguousarray(frames)[np.newaxis]
    assert origine_img.dtype == width.pix
    nt = origine_img.shape[1]
    if nt < 2:
        raise ValueError("the reference crashes for nt < 2 (compress.py:267)")
    X_test = origine_img.astype(np.float32) / width.pixmax              # :138
    X_test_pad = co.data_padding(X_test)                                # :176 (float64)
    key_frame = np.zeros(origine_img.shape, dtype=width.pix)            # :183
    PRE = int(p)
    windows = []
    t_pred = 0.0

    def predict2(frame):
        nonlocal t_pred
        x = np.stack([frame, np.zeros(frame.shape)], axis=0)[np.newaxis]   # :224-226
        t0 = tic()
        out = predictor.predict(x, 10)                                  # :227
        t_pred += tic() - t0
        return out

    X_hat = None
    if PRE:
        stack = []
        for w_idx in range(PRE):                                        # :189-202
            key_frame[0, w_idx] = origine_img[0, w_idx]
            X_hat = predict2(X_test_pad[0, w_idx])
            stack.append(X_hat[0, 0])
        windows.append((0, stack))                                      # :205-206
        cur_first, cur_preds = PRE, [X_hat[0, 0]]                       # :207-211
    key_idx = PRE + 1
    stop_point = 0
    idx = PRE + 1
    while idx < nt:                                                     # :217
        if idx == key_idx:
            inp = X_test_pad[0, idx - 1]                                # :219
            key_frame[0, idx - 1] = origine_img[0, idx - 1]             # :220
        else:
            inp = cur_preds[-1]                                         # :222
        X_hat = predict2(inp)
        pred = X_hat[0, 1]                                              # :229
        if idx == 1:
            cur_first, cur_preds = 0, [X_hat[0, 0], pred]               # :235-240
        else:
            cur_preds.append(pred)                                      # :242-243
        if idx >= key_idx:                                              # :245-246
            ps = np.stack(cur_preds[1:])[np.newaxis]
            stop_point = np.mean((X_test_pad[:, key_idx:idx + 1] - ps) ** 2)
        if (threshold is not None and stop_point > threshold) or \
                (window is not None and (idx - PRE) % window == 0):     # :249
            windows.append((cur_first, cur_preds[:-1]))                 # :251-254
            cur_first, cur_preds = idx, [X_hat[0, 0]]                   # :256-259
            if idx == nt - 1:
                key_frame[0, idx] = origine_img[0, idx]                 # :261
                cur_preds[0] = X_hat[0, 1]                              # :262
            key_idx = idx + 1
            stop_point = 0
        idx += 1
    windows.append((cur_first, cur_preds))                              # :267-268
    T["predict"] = t_pred
    out = encode_windows(origine_img, windows, PRE, mode, bound, entropy, width, T)
    out["key_plane"] = key_frame.ravel()
    out["keys"] = [f for f in range(nt) if key_frame[0, f].any()]
    out["Hp"], out["Wp"] = X_test_pad.shape[2], X_test_pad.shape[3]
    return out


def parse_payload(data, width=W16):
    """decompress.py:103-113,203-221 with the v2 trailer."""
    data = np.asarray(data, dtype=width.code)
    if width is W16:
        assert (int(data[-1]) & 0xffffffff) == MAGIC and int(data[-2]) == VERSION and int(data[-3]) == 16
        data = data[:-3]
    warm_up = int(data[-1])                                             # :106
    data = data[:-1]
    shape = tuple(int(v) for v in data[-5:])                            # :111
    data = data[:-5]
    table_len = int(data[-1])                                           # :204
    if table_len == -1:
        return data[:-1], None, shape, warm_up                          # :207
    table_start = -table_len - 1
    return data[:table_start], data[table_start:-1].copy(), shape, warm_up   # :209-221


def decompress_arrays(key_plane, payload, predictor, width=W16, timers=None):
    """decompress.py:94-256,269 on arrays -> frames [nt,H,W,C] in the sample type (+ info dict)."""
    T = timers if timers is not None else {}
    tic = time.perf_counter
    body, table, shape, warm_up = parse_payload(payload, width)
    X_test = np.asarray(key_plane, dtype=width.pix).reshape(shape)      # :94,115
    X_test = X_test / width.pixmax                                      # :117 (float64; equals the float32 quotient
    X_test = X_test.astype(np.float32).astype(np.float64)               #  of compress.py:138 once keras casts to floatx)
    X_test_pad = co.data_padding(X_test)                                # :120
    key_frame_check = [i for i in range(X_test_pad.shape[1]) if not np.all(X_test_pad[0, i] == 0)]  # :123-127
    key_frame_check.append(X_test_pad.shape[1])                         # :129
    n_calls = 0
    t_pred = 0.0

    def pred(x):
        nonlocal n_calls, t_pred
        t0 = tic()
        out = predictor.predict(x, 10)
        t_pred += tic() - t0
        n_calls += 1
        return out

    result_list = []
    warm_up_frame = pred(X_test_pad[0, 0][np.newaxis, np.newaxis])      # :141-143
    for _ in range(warm_up):
        result_list.append(warm_up_frame)                               # :144-145
    for idx in range(warm_up, len(key_frame_check[warm_up:]) + warm_up - 1):   # :147
        for predict_idx in range(key_frame_check[idx], key_frame_check[idx + 1]):
            if predict_idx == key_frame_check[idx]:
                one = X_test_pad[0, predict_idx][np.newaxis, np.newaxis]
                result_list.append(one)                                 # :156-158 (the discarded predict of :150-154 skipped)
            elif predict_idx == key_frame_check[idx] + 1:
                one = X_test_pad[0, predict_idx - 1][np.newaxis, np.newaxis]
                X_hat = pred(np.concatenate([one, np.zeros(one.shape)], axis=1))   # :161-165
                result_list.append(X_hat[0, 1][np.newaxis, np.newaxis])            # :167-169
            else:
                one = result_list[-1]
                X_hat = pred(np.concatenate([one, np.zeros(one.shape)], axis=1))   # :172-175
                result_list.append(X_hat[0, 1][np.newaxis, np.newaxis])            # :177-179
    T["predict"] = t_pred
    X_hat_flat = np.concatenate(result_list, axis=1).astype(np.float64)  # :182-184
    X_hat_flat[0, 0] = X_test_pad[0, 0]                                 # :186
    X_hat_no_pad = X_hat_flat[:, :, :shape[2], :shape[3]]               # :189
    if table is not None:
        body = replacing_decode(body, table)                            # :229
        body = np.subtract(width.code.type(width.offset), body)         # :236
    x = delta_decode(body).reshape(shape)                               # :240-245
    # :252 multiplies the float64 copy of the float32 prediction; the encoder (compress.py:307) truncated the FLOAT32
    # product.  The two truncations agree for every float32 in [0, 1] at 255 (exhaustively checked) but not
    # at 65535, so the extension states the decoder with the encoder's own float32 product -- the only choice that
    # keeps the round trip lossless.
    pred_levels = np.multiply(X_hat_no_pad.astype(np.float32), np.float32(width.pixmax)).astype(np.int64) \
        if width is W16 else None
    key_mask = np.zeros(shape[1], bool)
    key_mask[[k for k in key_frame_check[:-1] if k >= warm_up]] = True
    key_mask[0] = True
    if width is W16:
        dec = pred_levels.astype(np.float64)
        kf = np.asarray(key_plane, dtype=width.pix).reshape(shape).astype(np.float64)
        dec[:, key_mask] = kf[:, key_mask]
    else:
        dec = X_hat_no_pad * width.pixmax                               # :252
    dec = dec - x                                                       # :253
    dec = np.where(dec > width.pixmax, width.pixmax, dec)               # :255
    dec = np.where(dec < 0, 0, dec)                                     # :256
    out = dec.astype(width.pix)[0]                                      # :269
    return out, {"keys": key_frame_check[:-1], "n_predict_calls": n_calls, "p": warm_up, "shape": shape, "x": x.ravel()}
