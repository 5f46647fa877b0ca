"""GPU codec kernels vs the oracle, GIVEN IDENTICAL PREDICTIONS: the int16 stream, the table, the key plane and
the decoded frames must be bit-exact.  All calls go through the C ABI (tezip_b200.ops)."""
import numpy as np
import pytest

from helpers import TINY, oracle_net, pool_from_oracle
from tezip_b200 import synth

pytestmark = pytest.mark.gpu

CASES = [
    # nt, H, W, p, window, threshold, mode, bound, entropy
    (12, 20, 28, 0, 5, None, "abs", [0.0], True),
    (12, 20, 28, 2, 5, None, "abs", [2.0], True),
    (12, 20, 28, 0, 5, None, "abs", [2.55], True),
    (12, 20, 28, 1, 4, None, "absrel", [3.0, 0.1], False),
    (12, 20, 28, 0, 5, None, "pwrel", [0.03], True),
    (12, 20, 28, 0, 4, None, "rel", [0.02], True),
    (11, 16, 24, 0, 5, None, "abs", [0.0], False),     # rowlen % 8 == 0: vector paths
    (11, 16, 24, 3, 3, None, "abs", [1.0], True),
    (9, 16, 24, 0, None, 0.12, "abs", [0.0], True),    # DWP
    (12, 128, 160, 0, 5, None, "abs", [2.0], True),    # the BASELINE frame shape: 20480-sample planes, 20 tiles each
    (12, 128, 160, 1, 5, None, "abs", [0.0], True),
    (7, 128, 160, 0, 3, None, "rel", [0.02], True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(str(v) for v in c))
def test_codec_bit_exact_given_oracle_predictions(cuda_lib, case):
    import torch
    from oracle import codec_oracle as co
    from tezip_b200 import codec, ops
    nt, H, W, p, window, thr, mode, bound, entropy = case
    net, _ws = oracle_net(TINY)
    frames = synth.make_frames(nt, H, W, 3, seed=3)
    r = co.compress_arrays(frames, net, p, window, thr, mode, bound, entropy)
    dev = torch.device("cuda", 0)
    pool, pred_slot, apply_eb = pool_from_oracle(r, dev)
    fr = torch.from_numpy(frames).to(dev)
    enc = codec.encode_with_pool(fr, pool, pred_slot, apply_eb, r["keys"], p, mode, bound, entropy, keep_x=True)
    assert np.array_equal(enc.key_plane.cpu().numpy().ravel(), r["key_plane"])
    assert np.array_equal(enc.x.cpu().numpy().ravel(), r["x"])
    if entropy:
        assert np.array_equal(enc.table, r["table"])
    assert np.array_equal(enc.payload(), r["payload"])
    # decoder kernels on the oracle's stream with the oracle's predictions
    body, table, shape, pp = codec.parse_payload(r["payload"])
    ref_out, _info = co.decompress_arrays(r["key_plane"], r["payload"], net)
    slot = torch.from_numpy(pred_slot).to(dev)
    if pp > 1:
        pass  # warm-up frames 1..p-1 already point at their P0 copy in the pool
    lut = torch.from_numpy(ops.decode_lut(table)).to(dev) if table is not None else None
    out, x = ops.reconstruct(torch.from_numpy(np.ascontiguousarray(body)).to(dev), (nt, H, W, 3), pool.shape[1],
                             pool.shape[2], len(table) if table is not None else -1, lut, pool, slot,
                             enc.key_plane, want_x=True)
    assert np.array_equal(x.cpu().numpy(), r["x"])
    assert np.array_equal(out.cpu().numpy(), ref_out)


def test_fused_lossless_equals_unfused(cuda_lib):
    import torch
    from oracle import codec_oracle as co
    from tezip_b200 import codec
    net, _ = oracle_net(TINY)
    for (H, W) in ((20, 28), (16, 24)):
        frames = synth.make_frames(10, H, W, 3, seed=5)
        r = co.compress_arrays(frames, net, 0, 4, None, "abs", [0.0], True)
        dev = torch.device("cuda", 0)
        pool, pred_slot, apply_eb = pool_from_oracle(r, dev)
        fr = torch.from_numpy(frames).to(dev)
        a = codec.encode_with_pool(fr, pool, pred_slot, apply_eb, r["keys"], 0, "abs", [0.0], True, keep_x=False)
        b = codec.encode_with_pool(fr, pool, pred_slot, apply_eb, r["keys"], 0, "abs", [0.0], True, keep_x=True)
        assert np.array_equal(a.payload(), b.payload())
        assert np.array_equal(a.payload(), r["payload"])


def test_error_bound_random_planes(cuda_lib):
    """error_bound kernel vs the C oracle on random planes, every mode, awkward bounds."""
    import torch
    from oracle import codec_oracle as co
    from tezip_b200 import ops
    rng = np.random.default_rng(11)
    nt, H, W, C = 6, 12, 20, 3
    frames = rng.integers(0, 256, size=(nt, H, W, C), dtype=np.uint8)
    dev = torch.device("cuda", 0)
    for mode, bound in (("abs", [0.5]), ("abs", [1.0]), ("abs", [2.55]), ("abs", [7.9]), ("abs", [-3.0]),
                        ("rel", [0.013]), ("absrel", [4.0, 0.02]), ("absrel", [1.5, 0.5]), ("pwrel", [0.1]),
                        ("pwrel", [0.017])):
        d = rng.integers(-40, 41, size=(nt, H, W, C)).astype(np.int64)
        ref = d.copy()
        ref0 = np.ascontiguousarray(ref[None])
        # oracle works window-wise on frames >= 1: emulate with one window holding all frames
        co.error_bound_frames(frames.astype(np.int64), ref, mode, bound)
        x = torch.from_numpy(d.astype(np.int16)).to(dev)
        apply = np.ones(nt, np.uint8)
        apply[0] = 0
        ops.error_bound(torch.from_numpy(frames).to(dev), x, torch.from_numpy(apply).to(dev), mode, bound)
        assert np.array_equal(x.cpu().numpy().astype(np.int64), ref), (mode, bound)


@pytest.mark.parametrize("hw", [(1, 1), (3, 11), (64, 32), (41, 50), (70, 93), (128, 160)],
                         ids=lambda hw: "%dx%d" % hw)
def test_error_bound_tiled_planes(cuda_lib, hw):
    """The tiled error-bound kernel (2048-element tiles, 32-element chunks) vs the C oracle on planes that span
    zero, one and many tiles, with short segments (noise), long segments (flat + spikes, random walks: segments
    crossing chunk and tile boundaries), constant planes and the int16 extremes."""
    import torch
    from oracle import codec_oracle as co
    from tezip_b200 import ops
    H, W = hw
    C, nt = 3, 7
    rng = np.random.default_rng(H * 1000 + W)
    frames = rng.integers(0, 256, size=(nt, H, W, C), dtype=np.uint8)
    dev = torch.device("cuda", 0)
    n = H * W

    def planes(kind):
        if kind == "noise":
            return rng.integers(-3, 4, size=(nt, H, W, C))
        if kind == "wide":
            return rng.integers(-300, 301, size=(nt, H, W, C))
        if kind == "flat":
            d = np.zeros((nt, n, C), np.int64)
            for f in range(nt):
                for c in range(C):
                    k = max(1, n // 700)
                    d[f, rng.integers(0, n, size=k), c] = rng.integers(-30, 31, size=k)
            return d.reshape(nt, H, W, C)
        if kind == "walk":
            return np.cumsum(rng.integers(-1, 2, size=(nt, n, C)), axis=1).reshape(nt, H, W, C)
        if kind == "const":
            return np.full((nt, H, W, C), 5)
        if kind == "extremes":
            return rng.choice(np.array([-32768, -32767, -1, 0, 1, 32766, 32767]), size=(nt, H, W, C))
        raise ValueError(kind)

    for kind in ("noise", "wide", "flat", "walk", "const", "extremes"):
        for mode, bound in (("abs", [2.0]), ("abs", [0.5]), ("abs", [2.55]), ("abs", [40000.0]), ("rel", [0.013]),
                            ("absrel", [4.0, 0.02])):
            d = planes(kind).astype(np.int64)
            ref = d.copy()
            co.error_bound_frames(frames.astype(np.int64), ref, mode, bound)
            x = torch.from_numpy(d.astype(np.int16)).to(dev)
            apply = np.ones(nt, np.uint8)
            apply[0] = 0
            ops.error_bound(torch.from_numpy(frames).to(dev), x, torch.from_numpy(apply).to(dev), mode, bound)
            got = x.cpu().numpy().astype(np.int64)
            assert np.array_equal(got, ref.astype(np.int16).astype(np.int64)), (kind, mode, bound)


def test_window_sse_matches_numpy(cuda_lib):
    import torch
    from tezip_b200 import ops
    rng = np.random.default_rng(2)
    H, W, Hp, Wp = 20, 28, 24, 32
    frames = rng.integers(0, 256, size=(5, H, W, 3), dtype=np.uint8)
    pred = rng.random((3, Hp, Wp, 3), dtype=np.float32)
    idx = np.array([4, 0, 2], np.int32)
    dev = torch.device("cuda", 0)
    sse = ops.window_sse(torch.from_numpy(frames).to(dev), torch.from_numpy(idx).to(dev),
                         torch.from_numpy(pred).to(dev)).cpu().numpy()
    pad = np.zeros((3, Hp, Wp, 3))
    pad[:, :H, :W] = (frames[idx].astype(np.float32) / 255)
    ref = ((pad - pred.astype(np.float64)) ** 2).reshape(3, -1).sum(axis=1)
    assert np.allclose(sse, ref, rtol=1e-12, atol=0)


def test_device_table_matches_host(cuda_lib):
    """tz_build_table (table + symbol->rank LUT on the device) vs the host construction that mirrors
    compress.py:352-361 / :84-90: equal counts (tie -> ascending symbol), a single bin, every bin, and tables with a
    symbol inside the rank range (flagged, LUT then comes from the host)."""
    import torch
    from tezip_b200 import ops
    rng = np.random.default_rng(21)
    dev = torch.device("cuda", 0)
    cases = []
    for n, lo, hi in ((60, 1500, 1700), (1, 1600, 1601), (300, 1000, 2200), (4096, 0, 4096), (40, 100, 4000)):
        h = np.zeros(4096, np.int64)
        sy = rng.choice(np.arange(lo, hi), size=min(n, hi - lo), replace=False)
        h[sy] = rng.integers(1, 50, size=len(sy))            # many ties
        cases.append(h)
    h = np.zeros(4096, np.int64)
    h[[3, 1600, 1601, 1599]] = [5, 10 ** 12, 7, 7]           # symbol 3 lies inside the rank range [0, 4): collision
    cases.append(h)
    for h in cases:
        want_table = ops.build_table(h)
        table = torch.empty(4096, dtype=torch.int16, device=dev)
        lut = torch.empty(4096, dtype=torch.int16, device=dev)
        meta = torch.empty(2, dtype=torch.int32, device=dev)
        ops.build_table_device(torch.from_numpy(h).to(dev), table, lut, meta)
        n, bad = (int(v) for v in meta.cpu().numpy())
        assert n == len(want_table) and np.array_equal(table[:n].cpu().numpy(), want_table)
        collides = len(want_table) > 0 and int(want_table.min()) < len(want_table)
        assert bool(bad) == collides
        if not bad:
            assert np.array_equal(lut.cpu().numpy(), ops.encode_lut(want_table))


@pytest.mark.parametrize("H,W,C", [(128, 160, 3), (41, 56, 3), (1, 8, 3), (64, 32, 1), (33, 8, 1), (70, 96, 3),
                                   (96, 1024, 1)], ids=lambda v: str(v))
def test_fused_lossy_pass_equals_separate_kernels(cuda_lib, H, W, C):
    """tz_encode_lossy (residual + error bound + delta histogram in one data pass, speculative segment scan) against
    tz_residual + tz_error_bound + tz_delta_hist: x and the histogram must be identical, for short segments (noise),
    long segments (flat + spikes, random walks, ramps: segments crossing lane, tile and plane-tail boundaries),
    constant planes, window-start frames (x = 0) and frames without error bound (warm-up)."""
    import torch
    from tezip_b200 import ops
    from tezip_b200._lib import TZ_HIST_BINS
    dev = torch.device("cuda", 0)
    nt, n = 9, H * W
    rng = np.random.default_rng(H * 7919 + W * 13 + C)

    def planes(kind):
        if kind == "noise":
            return rng.integers(-3, 4, size=(nt, n, C))
        if kind == "wide":
            return rng.integers(-100, 101, size=(nt, n, C))
        if kind == "flat":
            d = np.zeros((nt, n, C), np.int64)
            for f in range(nt):
                for c in range(C):
                    k = max(1, n // 700)
                    d[f, rng.integers(0, n, size=k), c] = rng.integers(-30, 31, size=k)
            return d
        if kind == "walk":
            return np.clip(np.cumsum(rng.integers(-1, 2, size=(nt, n, C)), axis=1), -100, 100)
        if kind == "ramp":
            return np.broadcast_to(((np.arange(n) // 37) % 60 - 30)[None, :, None], (nt, n, C)).copy()
        return np.full((nt, n, C), 5)

    for kind in ("noise", "wide", "flat", "walk", "ramp", "const"):
        d = planes(kind).reshape(nt, H, W, C)
        a = rng.integers(100, 156, size=(nt, H, W, C))
        frames = torch.from_numpy(a.astype(np.uint8)).to(dev)
        pool = torch.from_numpy(((d + a + 0.5) / 255.0).astype(np.float32)).to(dev)      # trunc(pool * 255) = d + a
        slot_np = np.arange(nt, dtype=np.int32)
        slot_np[[0, 5]] = -1                                # window starts: x = 0
        apply_np = (slot_np >= 0).astype(np.uint8)
        apply_np[1] = 0                                     # a warm-up frame: residual, no error bound
        slot, apply = torch.from_numpy(slot_np).to(dev), torch.from_numpy(apply_np).to(dev)
        for mode, bound in (("abs", [2.0]), ("abs", [0.5]), ("abs", [2.55]), ("abs", [7.0]), ("abs", [40000.0]),
                            ("rel", [0.013]), ("absrel", [4.0, 0.02]), ("abs", [1.0000001])):
            x_ref = ops.residual(frames, pool, slot)
            assert np.array_equal(x_ref.cpu().numpy()[slot_np >= 0], (d[slot_np >= 0]).astype(np.int16))
            ops.error_bound(frames, x_ref, apply, mode, bound)
            h_ref = torch.zeros(TZ_HIST_BINS + 1, dtype=torch.int64, device=dev)
            ops.finding_difference_hist(x_ref, h_ref[:-1], h_ref[-1:])
            h = torch.zeros(TZ_HIST_BINS + 2, dtype=torch.int64, device=dev)
            x = ops.encode_lossy(frames, pool, slot, apply, mode, bound, h[:TZ_HIST_BINS], h[TZ_HIST_BINS:-1],
                                 h[-1:].view(torch.int32))
            assert torch.equal(x, x_ref), (kind, mode, bound)
            assert torch.equal(h[:TZ_HIST_BINS + 1], h_ref), (kind, mode, bound)
            # has_prev = 3: everything but the first symbol of the stream
            h3 = torch.zeros(TZ_HIST_BINS + 2, dtype=torch.int64, device=dev)
            ops.encode_lossy(frames, pool, slot, apply, mode, bound, h3[:TZ_HIST_BINS], h3[TZ_HIST_BINS:-1],
                             h3[-1:].view(torch.int32), has_prev=3)
            assert int(h3[:TZ_HIST_BINS].sum()) == x.numel() - 1
            ops.finding_difference_hist(x.view(-1)[:1], h3[:TZ_HIST_BINS], h3[TZ_HIST_BINS:-1], 1, 7)
            hp = torch.zeros(TZ_HIST_BINS + 1, dtype=torch.int64, device=dev)
            ops.finding_difference_hist(x_ref, hp[:-1], hp[-1:], 1, 7)
            assert torch.equal(h3[:TZ_HIST_BINS + 1], hp)


def _np_delta(x, has_prev, prev_x):
    """compress.py:73-77 in int16 arithmetic (y[0] = x[0], or prev_x - x[0] on a shard)."""
    x = x.astype(np.int16)
    y = np.empty_like(x)
    y[1:] = (x[:-1].astype(np.int32) - x[1:].astype(np.int32)).astype(np.int16)
    y[0] = x[0] if not has_prev else np.int16(np.int32(prev_x) - np.int32(x[0]))
    return y


@pytest.mark.parametrize("n", [1, 7, 8, 255, 1024, 1032, 4095, 4096, 8 * 1024 + 8, 300001])
def test_stream_kernels_chunk_edges(cuda_lib, n):
    """The warp-chunk stream kernels (1024 elements per warp, four rounds of 256) on a materialised x: lengths around
    the round / chunk / block boundaries, every has_prev mode (0 = start of the stream, 1 = shard with a halo element,
    2 = a chunk of a longer stream), against numpy -- histogram, rank map, raw delta stream, and the decoder's
    prefix-sum inverse of it."""
    import torch
    from tezip_b200 import ops, _lib
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(n)
    full = rng.randint(-40, 41, size=n + 8).astype(np.int16)
    full[rng.rand(n + 8) < 0.01] = 255            # a few large jumps
    xt_full = torch.from_numpy(full).to(dev)      # (chunks of a longer stream start at multiples of 8 elements)
    for has_prev, prev_x, x_np, x_t in ((0, 0, full[8:], xt_full[8:].clone()), (1, -17, full[8:], xt_full[8:].clone()),
                                        (2, int(full[7]), full[8:], xt_full[8:])):
        y = _np_delta(x_np, has_prev != 0, prev_x)
        sym = (1600 - y.astype(np.int32)).astype(np.int16)
        hist = torch.zeros(_lib.TZ_HIST_BINS, dtype=torch.int64, device=dev)
        ovf = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.finding_difference_hist(x_t, hist, ovf, has_prev, prev_x)
        assert np.array_equal(hist.cpu().numpy(), np.bincount(sym.astype(np.int64), minlength=_lib.TZ_HIST_BINS)), \
            (n, has_prev)
        assert int(ovf[0]) == 0
        table = ops.build_table(hist.cpu().numpy())
        lut = torch.from_numpy(ops.encode_lut(table)).to(dev)
        ranks = ops.finding_difference_rank(x_t, lut, has_prev=has_prev, prev_x=prev_x).cpu().numpy()
        pos = {int(s): i for i, s in enumerate(table)}
        assert np.array_equal(ranks, np.array([pos[int(s)] for s in sym], np.int16)), (n, has_prev)
        raw = ops.finding_difference_rank(x_t, None, has_prev=has_prev, prev_x=prev_x).cpu().numpy()
        assert np.array_equal(raw, y), (n, has_prev)


@pytest.mark.parametrize("H,W,C,nt", [(8, 8, 1, 70), (4, 8, 3, 41), (16, 24, 3, 9), (5, 7, 3, 30), (128, 160, 3, 3)])
def test_stream_kernels_many_small_frames(cuda_lib, H, W, C, nt):
    """Fused lossless passes and the decoder where a 1024-element chunk spans many frames (tiny frames: the in-frame
    offset wraps several times inside one chunk), rows that are / are not multiples of 8 samples, key frames inside
    the stream: the two passes equal the unfused kernels on the materialised residual, and the decoder returns the
    frames and x."""
    import torch
    from tezip_b200 import ops, _lib
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(H * W + nt)
    frames = rng.randint(0, 256, size=(nt, H, W, C)).astype(np.uint8)
    pool_np = rng.rand(nt + 1, H, W, C).astype(np.float32)
    slot_np = np.arange(1, nt + 1, dtype=np.int32)
    slot_np[::5] = -1                                   # window starts: x = 0, decoded from the key plane
    fr, pool, slot = (torch.from_numpy(a).to(dev) for a in (frames, pool_np, slot_np))
    x = ops.residual(fr, pool, slot)
    x_np = x.cpu().numpy().reshape(-1)
    q = (pool_np[np.maximum(slot_np, 0)] * np.float32(255)).astype(np.int64)      # trunc toward zero (values >= 0)
    ref = np.where((slot_np >= 0)[:, None, None, None], q - frames, 0).reshape(-1)
    assert np.array_equal(x_np, ref.astype(np.int16))
    hist_a = torch.zeros(_lib.TZ_HIST_BINS, dtype=torch.int64, device=dev)
    hist_b = torch.zeros_like(hist_a)
    ovf = torch.zeros(1, dtype=torch.int64, device=dev)
    ops.finding_difference_hist(x, hist_a, ovf)
    ops.encode_lossless(fr, pool, slot, 0, hist=hist_b, overflow=ovf)
    assert torch.equal(hist_a, hist_b) and int(ovf[0]) == 0
    table = ops.build_table(hist_a.cpu().numpy())
    lut = torch.from_numpy(ops.encode_lut(table)).to(dev)
    body_a = ops.finding_difference_rank(x, lut)
    body_b = torch.empty_like(body_a)
    ops.encode_lossless(fr, pool, slot, 1, lut=lut, out=body_b)
    assert torch.equal(body_a, body_b)
    key_plane = torch.from_numpy(np.where((slot_np < 0)[:, None, None, None], frames, 0).astype(np.uint8)).to(dev)
    lut_d = torch.from_numpy(ops.decode_lut(table)).to(dev)
    out, x_back = ops.reconstruct(body_a, (nt, H, W, C), H, W, len(table), lut_d, pool, slot, key_plane, want_x=True)
    assert np.array_equal(x_back.cpu().numpy(), x_np)
    assert np.array_equal(out.cpu().numpy(), frames)    # lossless: P - (P - frame), and the key frames verbatim
