"""ORACLE (test infrastructure only) -- CPU restatement of the reference's compress/decompress hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
numpy for the array arithmetic, oracle/codec_oracle.c (gcc) for the per-sample loops, pure-Python loops
kept only as the small-case cross-check (`error_bound_py`).  Every function cites the reference lines
(/root/reference/src/...) it follows.

PINNING: the reference ships no tests, fixtures or golden vectors.  This restatement is
pinned instead by (1) running the UNMODIFIED reference compress.py/decompress.py under oracle/refharness.py
in the build container and asserting byte-identical key_frame.dat / entropy.dat payloads and decoded frames
(tests/test_oracle_vs_reference.py; the same runs produced tests/golden/*.npz via tests/golden/make_golden.py)
and (2) the design-doc worked examples with the "code wins" corrections (tests/test_known_answers.py).
The PredNet arithmetic itself (TensorFlow 1.15 / cuDNN 7.6.5) is a third-party dependency that is absent
here: that part is "parity unpinned" and is restated in oracle/prednet_oracle.py from Keras semantics.
"""
import ctypes
import time

import numpy as np

from . import build as _build

MODES = {"abs": 0, "rel": 1, "absrel": 2, "pwrel": 3}
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        L.tzo_error_bound_plane.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
                                            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        L.tzo_delta_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        L.tzo_delta_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        L.tzo_replace.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_int]
        L.tzo_residual_frame.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_long] * 4
        for f in (L.tzo_error_bound_plane, L.tzo_delta_encode, L.tzo_delta_decode, L.tzo_replace,
                  L.tzo_residual_frame):
            f.restype = None
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------------ padding
def padding_size(num):
    """data_utils.py:103-107."""
    if num % 8 == 0:
        return num
    return (int(num / 8) + 1) * 8


def data_padding(X):
    """data_utils.py:77-92: zero-pad bottom/right to a multiple of 8; result is float64."""
    H, W = X.shape[2], X.shape[3]
    out = np.zeros((X.shape[0], X.shape[1], padding_size(H), padding_size(W), X.shape[4]))
    out[:, :, :H, :W] = X
    return out


# ------------------------------------------------------------------------------------------------ error_bound
def error_bound_py(origine, diff, mode, value):
    """compress.py:23-70 verbatim in structure (pure Python loop; small cases only)."""
    if value[0] == 0:
        return diff
    Bf = origine.flatten()
    Df = diff.flatten()
    if mode == "abs":
        E = np.abs(value[0])
    elif mode == "rel":
        E = (Bf.max() - Bf.min()) * value[0]
    elif mode == "absrel":
        if value[1] == 0:
            return diff
        abs_value = np.abs(value[0])
        rel_value = (Bf.max() - Bf.min()) * value[1]
        E = abs_value if abs_value < rel_value else rel_value
    elif mode == "pwrel":
        E = Bf * value[0]
    Du = Df + E
    Dl = Df - E
    u = float(np.inf)
    l = -u
    head = 0
    for i in range(len(Df)):
        if min((u, Du[i])) - max((l, Dl[i])) < 0.0:
            Df[head:i] = (u + l) / 2
            u = float(np.inf)
            l = -u
            head = i
        if Du[i] < u:
            u = Du[i]
        if l < Dl[i]:
            l = Dl[i]
    Df[head:len(Df)] = (u + l) / 2
    return Df.reshape(diff.shape)


def error_bound_frames(X_int, diff, mode, value):
    """compress.py:315-319 for one window: frames >= 1, each channel plane, in place on `diff`.
    X_int / diff: int64 [n,H,W,C] C-contiguous."""
    n, H, W, C = diff.shape
    if value[0] == 0:
        return
    b1 = float(value[1]) if len(value) > 1 else 0.0
    scratch = np.empty(2 * H * W, np.float64)
    L = lib()
    for f in range(1, n):
        for c in range(C):
            L.tzo_error_bound_plane(X_int[f, :, :, c].ctypes.data, diff[f, :, :, c].ctypes.data,
                                    H * W, C, MODES[mode], float(value[0]), b1, scratch.ctypes.data)


# ------------------------------------------------------------------------------------------------ delta / table
def delta_encode(x):
    """compress.py:73-77."""
    x = np.ascontiguousarray(x, np.int16).ravel()
    y = np.empty_like(x)
    lib().tzo_delta_encode(x.ctypes.data, y.ctypes.data, x.size)
    return y


def delta_decode(y):
    """decompress.py:22-29 (the reference's pure-Python per-element loop)."""
    y = np.ascontiguousarray(y, np.int16).ravel()
    x = np.empty_like(y)
    lib().tzo_delta_decode(y.ctypes.data, x.ctypes.data, y.size)
    return x


def build_table(s):
    """compress.py:352-361: symbols with count > 0, by count descending; Python's stable sort with
    reverse=True keeps equal counts in ascending-symbol order."""
    y_elem = np.bincount(s)
    ii = np.nonzero(y_elem)[0]
    d = list(zip(ii, y_elem[ii]))
    d.sort(key=lambda e: e[1], reverse=True)
    return np.array([k for k, _ in d], dtype="int16")


def replacing_encode(arr, table):
    """compress.py:84-90 (sequential where() passes)."""
    out = np.ascontiguousarray(arr, np.int16).copy()
    table = np.ascontiguousarray(table, np.int16)
    lib().tzo_replace(out.ctypes.data, out.size, table.ctypes.data, table.size, 0)
    return out


def replacing_decode(arr, table):
    """decompress.py:31-36."""
    out = np.ascontiguousarray(arr, np.int16).copy()
    table = np.ascontiguousarray(table, np.int16)
    lib().tzo_replace(out.ctypes.data, out.size, table.ctypes.data, table.size, 1)
    return out


# ------------------------------------------------------------------------------------------------ compress
def compress_arrays(frames, predictor, p, window, threshold, mode, bound, entropy=True, timers=None):
    """compress.py:138-395 on arrays: frames u8 [nt,H,W,3] -> dict with
      key_plane  u8 [nt*H*W*3]           (compress.py:183,190,220,261,271-273)
      payload    int16                    (entropy.dat before zstd; compress.py:375-395)
      keys       list of key-frame indices
      windows    list of (first_frame, n_frames)
      preds      f32 [nt,Hp,Wp,3]         prediction paired with every frame (placeholder P0 at window starts)
      x, y       int16 [N]                quantised residuals / delta stream
    predictor: object with keras-like .predict(x[1,T,Hp,Wp,C]) -> [1,T,Hp,Wp,C] float32."""
    T = timers if timers is not None else {}
    tic = time.perf_counter
    origine_img = np.ascontiguousarray(frames)[np.newaxis]
    nt = origine_img.shape[1]
    if nt < 2:
        raise ValueError("the reference crashes for nt < 2 (compress.py:267)")
    X_test = origine_img.astype(np.float32) / 255                       # :138
    X_test_pad = data_padding(X_test)                                   # :176 (float64)
    key_frame = np.zeros(origine_img.shape, dtype="uint8")              # :183
    PRE = int(p)
    windows = []            # (first frame index, [pred arrays f32 Hp,Wp,C], n frames)
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

    out = encode_windows(origine_img, windows, PRE, mode, bound, entropy, T)
    out["key_plane"] = key_frame.ravel()
    out["keys"] = [f for f in range(nt) if key_frame[0, f].any()]   # what the decoder will see (decompress.py:123-127)
    out["Hp"], out["Wp"] = X_test_pad.shape[2], X_test_pad.shape[3]
    return out


def encode_windows(origine_img, windows, PRE, mode, bound, entropy=True, timers=None):
    """compress.py:289-395: residual, error bound, delta, table, rank map and trailer, GIVEN the windows
    [(first_frame, [prediction f32 [Hp,Wp,C] per frame])] the scheduler produced."""
    T = timers if timers is not None else {}
    tic = time.perf_counter
    X_test_shape = origine_img.shape
    H, W = X_test_shape[2], X_test_shape[3]
    difference_list = []
    t0 = tic()
    t_eb = 0.0
    for wi, (first, preds) in enumerate(windows):
        n = len(preds)
        origine_pick = origine_img[:, first:first + n] / 255            # :294
        predict_pick = np.stack(preds)[np.newaxis]                      # :295
        predict_pick_no_pad = predict_pick[:, :, :H, :W]                # :298
        X_hat_1 = np.multiply(predict_pick_no_pad, np.float32(255.0))   # :307 (float32 array * python float)
        assert X_hat_1.dtype == np.float32
        X_test_1 = np.multiply(origine_pick, 255.000)                   # :308
        X_test_1 = X_test_1.astype(int)                                 # :310
        X_hat_1 = X_hat_1.astype(int)                                   # :311
        difference = X_hat_1 - X_test_1                                 # :313
        difference[:, 0] = 0                                            # :314
        if not (PRE != 0 and wi == 0):                                  # :315
            t1 = tic()
            d0 = np.ascontiguousarray(difference[0])
            error_bound_frames(np.ascontiguousarray(X_test_1[0]), d0, mode, list(bound))   # :316-319
            difference = d0[np.newaxis]
            t_eb += tic() - t1
        difference_list.append(difference)
    T["residual"] = tic() - t0 - t_eb
    T["error_bound"] = t_eb

    difference_model = np.concatenate(difference_list, axis=1).astype("int16")   # :329-333
    x = difference_model.ravel().copy()
    t0 = tic()
    y = delta_encode(difference_model)                                  # :339-340
    T["finding_difference"] = tic() - t0
    result = y
    table = None
    if entropy:
        t0 = tic()
        s = np.subtract(np.int16(1600), y)                              # :348
        table = build_table(s)                                          # :352-361
        T["table_create"] = tic() - t0
        t0 = tic()
        result = replacing_encode(s, table)                             # :369
        T["replacing"] = tic() - t0
    t0 = tic()
    tail = []
    if entropy:
        tail += [int(v) for v in table] + [len(table)]                  # :383-385
    else:
        tail += [-1]                                                    # :387
    tail += [int(v) for v in X_test_shape] + [PRE]                      # :390-392
    payload = np.concatenate([result.astype(np.int64), np.array(tail, np.int64)]).astype(np.int16)   # :394
    T["pack"] = tic() - t0

    preds_full = np.concatenate([np.stack(w[1]) for w in windows], axis=0)
    return {"payload": payload, "windows": [(w[0], len(w[1])) for w in windows], "preds": preds_full,
            "x": x, "y": y, "table": table, "shape": tuple(X_test_shape)}


# ------------------------------------------------------------------------------------------------ decompress
def parse_payload(data):
    """decompress.py:103-113,203-221: returns (body int16, table or None, shape(5), p)."""
    data = np.asarray(data, dtype="int16")
    warm_up = int(data[-1])                                             # :106
    data = data[:-1]
    shape = tuple(int(v) for v in data[-5:])                            # :111
    data = data[:-5]
    table_len = int(data[-1])                                           # :204
    if table_len == -1:
        return data[:-1], None, shape, warm_up                          # :207
    table_start = -table_len - 1
    table = data[table_start:-1].copy()                                 # :209-212
    return data[:table_start], table, shape, warm_up                    # :221


def decompress_arrays(key_plane, payload, predictor, timers=None, replay_preds=None):
    """decompress.py:94-256,269 on arrays -> frames u8 [nt,H,W,3] (+ info dict).
    replay_preds (f32 [nt,Hp,Wp,C]): use these as the regenerated predictions instead of calling the predictor
    (golden-fixture tests: predictions recorded from the reference run)."""
    T = timers if timers is not None else {}
    tic = time.perf_counter
    body, table, shape, warm_up = parse_payload(payload)
    X_test = np.asarray(key_plane, dtype="uint8").reshape(shape)        # :94,115
    X_test = X_test / 255                                               # :117
    X_test_pad = data_padding(X_test)                                   # :120
    key_frame_check = [i for i in range(X_test_pad.shape[1]) if not np.all(X_test_pad[0, i] == 0)]  # :123-127
    key_frame_check.append(X_test_pad.shape[1])                         # :129
    t_pred = 0.0
    n_calls = 0

    frame_of_call = []

    def pred(x):
        nonlocal t_pred, n_calls
        t0 = tic()
        if replay_preds is not None:
            f = frame_of_call[-1]
            out = np.stack([replay_preds[0], replay_preds[f]])[np.newaxis][:, :x.shape[1] if x.shape[1] == 1 else 2]
        else:
            out = predictor.predict(x, 10)
        t_pred += tic() - t0
        n_calls += 1
        return out

    result_list = []
    frame_of_call.append(0)
    warm_up_frame = pred(X_test_pad[0, 0][np.newaxis, np.newaxis])      # :141-143 (one time step)
    for _ in range(warm_up):
        result_list.append(warm_up_frame)                               # :144-145
    for idx in range(warm_up, len(key_frame_check[warm_up:]) + warm_up - 1):   # :147
        for predict_idx in range(key_frame_check[idx], key_frame_check[idx + 1]):
            frame_of_call.append(predict_idx)
            if predict_idx == key_frame_check[idx]:
                one = X_test_pad[0, predict_idx][np.newaxis, np.newaxis]
                pred(np.concatenate([one, np.zeros(one.shape)], axis=1))         # :150-154 (discarded)
                result_list.append(one)                                          # :156-158
            elif predict_idx == key_frame_check[idx] + 1:
                one = X_test_pad[0, predict_idx - 1][np.newaxis, np.newaxis]
                X_hat = pred(np.concatenate([one, np.zeros(one.shape)], axis=1))  # :161-165
                result_list.append(X_hat[0, 1][np.newaxis, np.newaxis])          # :167-169
            else:
                one = result_list[-1]
                X_hat = pred(np.concatenate([one, np.zeros(one.shape)], axis=1))  # :172-175
                result_list.append(X_hat[0, 1][np.newaxis, np.newaxis])          # :177-179
    T["predict"] = t_pred
    X_hat_flat = np.concatenate(result_list, axis=1)                    # :182-184 (float64 by promotion)
    X_hat_flat = X_hat_flat.astype(np.float64)
    X_hat_flat[0, 0] = X_test_pad[0, 0]                                 # :186
    X_hat_no_pad = X_hat_flat[:, :, :X_test.shape[2], :X_test.shape[3]]  # :189
    t0 = tic()
    if table is not None:
        body = replacing_decode(body, table)                            # :229
        body = np.subtract(np.int16(1600), body)                        # :236
    T["replacing"] = tic() - t0
    t0 = tic()
    difference_first = delta_decode(body).reshape(shape)                # :240-245
    T["finding_difference"] = tic() - t0
    t0 = tic()
    dec = X_hat_no_pad * 255                                            # :252
    dec = dec - difference_first                                        # :253
    dec = np.where(dec > 255, 255, dec)                                 # :255
    dec = np.where(dec < 0, 0, dec)                                     # :256
    out = dec.astype("uint8")[0]                                        # :269
    T["reconstruct"] = tic() - t0
    return out, {"keys": key_frame_check[:-1], "n_predict_calls": n_calls, "p": warm_up, "shape": shape}
