"""The 16-bit extension's oracle (oracle/wide_oracle.py).  The reference cannot run 16-bit data, so the extension is
pinned by construction: the same functions with the reference's own constants (8-bit pixel maximum, int16 codes,
offset 1600) must reproduce oracle/codec_oracle.py -- itself pinned byte for byte against the unmodified reference and
its golden fixtures -- and with the widened constants they must round-trip exactly and respect the bound."""
import glob
import os

import numpy as np
import pytest

from oracle import codec_oracle as co, wide_oracle as wo
from oracle.prednet_oracle import PredNetOracle
from tezip_b200 import synth

MONO = (1, 16, 32, 64)
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("p,window,thr,mode,bound,entropy", [
    (0, 4, None, "abs", [0.0], True), (2, 5, None, "abs", [2.0], True), (1, 4, None, "absrel", [3.0, 0.1], False),
    (0, 5, None, "pwrel", [0.03], True), (0, None, 0.05, "rel", [0.02], True)])
def test_same_functions_with_reference_constants_equal_the_pinned_oracle(p, window, thr, mode, bound, entropy):
    stack = (3, 8, 16, 32)
    ws = synth.make_weights(stack, bias="uniform", seed=3)
    net = PredNetOracle(ws, stack, stack)
    frames = synth.make_frames(11, 12, 20, 3, seed=23)
    a = co.compress_arrays(frames, net, p, window, thr, mode, bound, entropy)
    b = wo.compress_arrays(frames, net, p, window, thr, mode, bound, entropy, width=wo.W8)
    assert np.array_equal(a["payload"], b["payload"]) and a["payload"].dtype == b["payload"].dtype
    assert np.array_equal(a["key_plane"], b["key_plane"]) and np.array_equal(a["x"], b["x"])
    out_a, _ = co.decompress_arrays(a["key_plane"], a["payload"], net)
    out_b, _ = wo.decompress_arrays(b["key_plane"], b["payload"], net, width=wo.W8)
    assert np.array_equal(out_a, out_b)


@pytest.mark.parametrize("path", [g for g in GOLD if "full_" not in g], ids=lambda p: os.path.basename(p)[:-4])
def test_reference_constants_reproduce_golden_payloads(path):
    """encode_windows with the reference's widths on the predictions the REFERENCE recorded == the reference's file."""
    from test_oracle_golden import load
    g = load(path)
    windows = [(int(f), [g["preds"][int(f) + i] for i in range(int(n))]) for f, n in g["windows"]]
    r = wo.encode_windows(g["frames"][np.newaxis], windows, g["p"], g["mode"], g["bound"], g["entropy"], width=wo.W8)
    assert np.array_equal(r["payload"], g["ref_payload"])


def _frames16(nt, H, W, seed):
    return synth.make_frames(nt, H, W, 1, seed=seed, dtype=np.uint16)


@pytest.mark.parametrize("p,window,thr,mode,bound,entropy,lim", [
    (0, 4, None, "abs", [0.0], True, 0), (1, 3, None, "abs", [0.0], False, 0), (0, 4, None, "abs", [300.0], True, 300),
    (2, 5, None, "abs", [77.5], True, 78), (0, 4, None, "rel", [0.01], True, None), (0, 3, None, "pwrel", [0.01], True, None),
    (0, None, 0.2, "abs", [0.0], True, 0)])
def test_wide_roundtrip(p, window, thr, mode, bound, entropy, lim):
    ws = synth.make_weights(MONO, bias="uniform", seed=7)
    net = PredNetOracle(ws, MONO, MONO)
    frames = _frames16(10, 20, 28, seed=5)
    r = wo.compress_arrays(frames, net, p, window, thr, mode, bound, entropy)
    assert r["payload"].dtype == np.int32 and r["key_plane"].dtype == np.uint16
    assert (int(r["payload"][-1]) & 0xffffffff) == wo.MAGIC
    if entropy:
        t = r["table"].astype(np.int64)
        assert t.min() >= wo.SYM_MIN and t.max() < wo.SYM_MIN + wo.NBINS and t.min() > len(t)
    out, info = wo.decompress_arrays(r["key_plane"], r["payload"], net)
    assert info["keys"] == r["keys"] and np.array_equal(info["x"], r["x"])
    err = np.abs(out.astype(np.int64) - frames.astype(np.int64))
    if lim is not None:
        assert err.max() <= lim
    elif mode == "rel":
        assert err.max() <= np.floor((int(frames.max()) - int(frames.min())) * bound[0]) + 1
    else:
        assert (err <= np.floor(frames.astype(np.float64) * bound[0]) + 1).all()
    if lim == 0:
        assert np.array_equal(out, frames)


def test_lookup_equals_the_sequential_where_passes():
    """compress.py:84-90 / decompress.py:31-36 literally (one where() pass per table entry) against the look-up."""
    rng = np.random.default_rng(3)
    y = rng.integers(-3000, 3001, 4000).astype(np.int32)
    s = np.int32(wo.OFFSET) - y
    table = wo.build_table(s, wo.W16)
    enc = wo.replacing_encode(s, table)
    lit = s.copy()
    for idx, num in enumerate(table):
        lit = np.where(lit == num, idx, lit)
    assert np.array_equal(enc, lit)
    dec = wo.replacing_decode(enc, table)
    lit2 = enc.copy()
    for idx, num in enumerate(table):
        lit2 = np.where(lit2 == idx, num, lit2)
    assert np.array_equal(dec, lit2) and np.array_equal(dec, s)
    assert np.array_equal(wo.delta_decode(wo.delta_encode(y)), y)


def test_host_payload_v2_matches_oracle():
    from tezip_b200 import codec, ops
    ws = synth.make_weights(MONO, bias="uniform", seed=7)
    net = PredNetOracle(ws, MONO, MONO)
    frames = _frames16(7, 16, 24, seed=9)
    for entropy in (True, False):
        r = wo.compress_arrays(frames, net, 1, 3, None, "abs", [0.0], entropy)
        body, table, shape, p = codec.parse_payload(r["payload"])
        assert body.dtype == np.int32 and shape == (1, 7, 16, 24, 1) and p == 1
        assert (table is None) == (not entropy)
        assert np.array_equal(codec.pack_payload(body, table, shape, p), r["payload"])
        assert codec.is_v2_payload(r["payload"]) and codec.is_v2_payload(r["payload"].tobytes() and np.frombuffer(r["payload"].tobytes(), np.uint8))
        if entropy:
            hist = np.bincount((np.int64(wo.OFFSET) - r["y"].astype(np.int64)) - wo.SYM_MIN, minlength=wo.NBINS)
            assert np.array_equal(ops.build_table16(hist), r["table"])
            lut = ops.decode_lut16(table)
            assert np.array_equal(lut[body], np.int32(wo.OFFSET) - r["y"])
    for path in GOLD[:3]:                                   # no 8-bit stream looks like a v2 one
        z = np.load(path)
        assert not codec.is_v2_payload(z["ref_payload"])
        b, t, sh, pp = codec.parse_payload(z["ref_payload"])
        assert b.dtype == np.int16 and sh[4] == 3
