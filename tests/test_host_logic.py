"""Host-side logic of tezip_b200 (no GPU): schedules, payload format, LUTs, container, CLI grammar, C ABI exports."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import codec_oracle as co
from tezip_b200 import codec, container, ops, synth, dist as tzdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _ToyNet:
    """predictor with the oracle's interface whose 'prediction' is a cheap function of the input"""
    def predict(self, x, batch_size=None):
        x = np.asarray(x, np.float32)
        out = np.empty_like(x)
        out[:, 0] = 0.25
        if x.shape[1] > 1:
            out[:, 1] = np.clip(0.9 * x[:, 0] + 0.02, 0, 1)
        return out


@pytest.mark.parametrize("nt,p,window", [(12, 0, 5), (12, 2, 5), (11, 0, 5), (25, 3, 10), (7, 0, 1), (9, 1, 20)])
def test_swp_schedule_matches_oracle(nt, p, window):
    frames = np.random.default_rng(0).integers(1, 256, (nt, 8, 8, 3), dtype=np.uint8)
    r = co.compress_arrays(frames, _ToyNet(), p, window, None, "abs", [0.0], True)
    keys = codec.swp_keys(nt, p, window)
    assert keys == r["keys"]
    plan = codec.plan_from_keys(nt, p, keys)
    real = [w for w in r["windows"] if not (p and w[0] == 0)]
    assert plan.windows == real
    # every non-key frame gets exactly one slot; window starts get -1; warm-up frames 1..p-1 use P0 (slot 0)
    assert (plan.pred_slot[[w[0] for w in real]] == -1).all()
    used = plan.pred_slot[plan.pred_slot > 0]
    assert len(set(used.tolist())) == len(used) == nt - len(keys) and plan.n_slots == len(used) + 1
    if p > 1:
        assert (plan.pred_slot[1:p] == 0).all()
    assert plan.apply_eb.sum() == nt - len(keys)
    # lock-step: step k advances every window longer than k, from a contiguous prefix
    for k, (kidx, slot0, B) in enumerate(plan.steps, start=1):
        assert B == sum(1 for w in real if w[1] > k)
    with pytest.raises(Exception):
        codec.swp_keys(p + 1, p, window)


def test_payload_pack_parse_matches_oracle():
    frames = np.random.default_rng(1).integers(1, 256, (9, 8, 8, 3), dtype=np.uint8)
    for entropy in (True, False):
        r = co.compress_arrays(frames, _ToyNet(), 1, 4, None, "abs", [1.0], entropy)
        body, table, shape, p = codec.parse_payload(r["payload"])
        b2, t2, s2, p2 = co.parse_payload(r["payload"])
        assert np.array_equal(body, b2) and shape == s2 == (1, 9, 8, 8, 3) and p == p2 == 1
        assert (table is None) == (t2 is None) == (not entropy)
        if entropy:
            assert np.array_equal(table, t2)
        assert np.array_equal(codec.pack_payload(body, table, shape, p), r["payload"])


def test_payload_trailer_limits():
    """The trailer stores nt, H, W and p as int16 (compress.py:390-394): larger values are refused, not wrapped."""
    body = np.zeros(4, np.int16)
    codec.pack_payload(body, None, (1, 32767, 1, 1, 3), 0)
    for shape, p in (((1, 32768, 8, 8, 3), 0), ((1, 10, 40000, 8, 3), 0), ((1, 10, 8, 8, 3), 40000)):
        with pytest.raises(Exception):
            codec.pack_payload(body, None, shape, p)


def test_luts_match_reference_replacing():
    rng = np.random.default_rng(2)
    y = rng.integers(-40, 41, 5000).astype(np.int16)
    s = (1600 - y).astype(np.int16)
    hist = np.bincount(s, minlength=4096)
    table = ops.build_table(hist)
    assert np.array_equal(table, co.build_table(s))
    enc = ops.encode_lut(table)[s]
    assert np.array_equal(enc, co.replacing_encode(s, table))
    dec = ops.decode_lut(table)[enc]
    assert np.array_equal(dec, co.replacing_decode(enc, table)) and np.array_equal(dec, s)
    # a colliding table (symbols inside the rank range) goes through the same sequential where() passes
    t2 = np.array([3, 1, 0, 2], np.int16)
    arr = np.array([0, 1, 2, 3, 7], np.int16)
    assert np.array_equal(ops.encode_lut(t2)[arr], co.replacing_encode(arr, t2))
    assert np.array_equal(ops.decode_lut(t2)[arr], co.replacing_decode(arr, t2))
    # random tables full of value/index collisions, over the whole symbol domain
    dom = np.arange(4096, dtype=np.int16)
    for trial in range(60):
        n = int(rng.integers(1, 300))
        hi = int(rng.choice([n, 2 * n, 4096]))
        t3 = rng.choice(hi, size=min(n, hi), replace=False).astype(np.int16)
        assert np.array_equal(ops.encode_lut(t3), co.replacing_encode(dom, t3)), trial
        assert np.array_equal(ops.decode_lut(t3), co.replacing_decode(dom, t3)), trial


def test_container_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    kp = rng.integers(0, 255, 3000, dtype=np.uint8)
    payload = rng.integers(-5, 2000, 4000).astype(np.int16)
    for workers in (0, 2):
        d = str(tmp_path / ("c%d" % workers))
        container.write_container(d, ["a.png", "b.png"], True, kp, payload, workers=workers)
        names, rgb, kp2, pl2 = container.read_container(d)
        assert names == ["a.png", "b.png"] and rgb and np.array_equal(kp, kp2) and np.array_equal(payload, pl2)
        # single zstd frame carrying its content size (what python-zstd's one-shot decompress needs)
        from oracle import refharness as rh
        raw = open(os.path.join(d, "entropy.dat"), "rb").read()
        assert rh.zstd_decompress(raw) == payload.tobytes()
        assert open(os.path.join(d, "filename.txt")).read() == "1\na.png\nb.png\n"


def test_shard_ranges_cover_and_align():
    for nt, p, w, world in ((1000, 0, 10, 8), (105, 2, 10, 4), (37, 0, 5, 2), (12, 1, 5, 8)):
        rs = tzdist.shard_ranges(nt, p, w, world)
        assert rs[0][0] == 0 and rs[-1][1] == nt
        for (a, b), (c, d) in zip(rs, rs[1:]):
            assert b == c
        keys = set(codec.swp_keys(nt, p, w))
        for r, (a, b) in enumerate(rs):
            if r > 0 and b > a:
                assert a in keys


def _cli(*argv):
    return subprocess.run([sys.executable, "-m", "tezip_b200.tezip", *argv], cwd=ROOT, capture_output=True, text=True)


def test_cli_flag_grammar():
    r = _cli("-c", "m", "i", "o", "-w", "5", "-m", "abs", "-b", "0")
    assert "Please specify the -p or --preprocess option!" in r.stdout
    r = _cli("-c", "m", "i", "o", "-p", "0", "-m", "abs", "-b", "0")
    assert "Please specify the window size(-w or --window) or MSE threshold(-t or --threshold) option!" in r.stdout
    r = _cli("-c", "m", "i", "o", "-p", "0", "-w", "5", "-t", "0.1", "-m", "abs", "-b", "0")
    assert "Please select only one of window size" in r.stdout
    r = _cli("-c", "m", "i", "o", "-p", "0", "-w", "5", "-m", "xyz", "-b", "0")
    assert "Please specify the -m or --mode correctly!" in r.stdout
    r = _cli("-c", "m", "i", "o", "-p", "0", "-w", "5", "-m", "absrel", "-b", "1")
    assert "enter two in -b or --bound" in r.stdout
    r = _cli("-c", "m", "i", "o", "-u", "m", "i", "o")
    assert "Please select only one of learn or compress or uncompress." in r.stdout
    r = _cli()
    assert "Please mode select!" in r.stdout
    r = _cli("-c", "m", "i", "o", "-p", "0", "-w", "5", "-m", "abs", "-b", "0", "-f")
    assert "not available" in r.stdout and r.returncode != 0
    h = _cli("-h").stdout
    for flag in ("-l", "-c", "-u", "-p", "-w", "-t", "-m", "-b", "-f", "-v", "-n"):
        assert flag in h


def test_cabi_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports every function include/tezip_b200.h declares; compute
    entry points fail loudly (no CPU fallback)."""
    import ctypes
    from tezip_b200 import build, _lib
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "tezip_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tz_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "python binding lacks %s" % name
    assert set(_lib.SIGNATURES) == declared
    assert lib.tz_abi_version() == _lib.TZ_ABI_VERSION == 2
    import torch
    if not torch.cuda.is_available():
        assert lib.tz_device_count() < 0 and b"cuda" in lib.tz_last_error().lower()
        cfg = _lib.PrednetConfig()
        cfg.n_layers, cfg.Hp, cfg.Wp, cfg.max_batch = 2, 8, 8, 1
        cfg.stack_sizes[0] = cfg.stack_sizes[1] = cfg.r_stack_sizes[0] = cfg.r_stack_sizes[1] = 3
        n = 2 * (6 * 2 - 1)
        ptrs = (ctypes.c_void_p * n)()
        sizes = (ctypes.c_longlong * n)()
        h = ctypes.c_void_p()
        assert lib.tz_prednet_create(ctypes.byref(cfg), ptrs, sizes, n, ctypes.byref(h)) != 0
        assert lib.tz_last_error()


def test_keras_weight_converter_orders_by_weight_names(tmp_path):
    """scripts/convert_keras_weights.py reads the datasets in the `weight_names` order of the one layer that owns
    weights (Keras' get_weights() order), from a weights-only file or the model_weights group of a full-model file;
    h5py is not in this image, so the file is stood in for by objects with the same attrs / [] shape."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("convert_keras_weights",
                                                  os.path.join(ROOT, "scripts", "convert_keras_weights.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)

    class Node(dict):
        def __init__(self, items=(), **attrs):
            super().__init__(items)
            self.attrs = attrs

    ws = synth.make_weights((3, 8, 16), bias="uniform", seed=5)
    names = ["pred_net_1/w%d:0" % i for i in range(len(ws))]
    order = np.random.default_rng(0).permutation(len(ws))                 # storage order differs from weight_names
    layer = Node({names[i]: ws[i] for i in order}, weight_names=[n.encode() for n in names])
    root = Node({"input_1": Node(weight_names=[]), "pred_net_1": layer}, layer_names=[b"input_1", b"pred_net_1"])
    got = conv.prednet_weights(root)
    assert len(got) == len(ws) and all(np.array_equal(a, b) for a, b in zip(got, ws))
    full = Node({"model_weights": root, "optimizer_weights": Node()})
    got = conv.prednet_weights(full)
    assert all(np.array_equal(a, b) for a, b in zip(got, ws))
    with pytest.raises(ValueError):
        conv.prednet_weights(Node({"x": Node()}))
    with pytest.raises(ValueError):                                       # two layers with weights: not this model
        conv.prednet_weights(Node({"a": layer, "b": layer}, layer_names=[b"a", b"b"]))


def test_shard_layouts_from_the_advisor_cases():
    """Round-1 advisor findings: a trailing one-frame window is a legal shard (its rank encodes a lone key frame);
    with fewer windows than ranks the decode ranges give rank 0 a real window and leave the TRAILING ranks empty."""
    from tezip_b200 import codec
    from tezip_b200.dist import shard_ranges, key_aligned_ranges
    assert shard_ranges(11, 0, 5, 3) == [(0, 5), (5, 10), (10, 11)]
    assert codec.swp_keys(1, 0, 5, shard=True) == [0]
    with pytest.raises(codec.TezipError):
        codec.swp_keys(1, 0, 5)
    plan = codec.plan_from_keys(1, 0, [0])
    assert plan.steps == [] and plan.n_slots == 1 and list(plan.pred_slot) == [-1]
    assert key_aligned_ranges([0, 1, 2, 7, 12], 15, 2, 4) == [(0, 7), (7, 12), (12, 15), (15, 15)]
    for keys, nt, p, world in (([0, 5, 10, 15], 20, 0, 2), ([0, 5, 10, 15, 20, 25, 30], 33, 0, 3), ([0, 1, 4], 9, 1, 8)):
        rs = key_aligned_ranges(keys, nt, p, world)
        assert rs[0][0] == 0 and rs[-1][1] == nt and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert rs[0][1] > p                                   # rank 0 owns a real window, not just warm-up frames
        assert all(a in keys for a, b in rs[1:] if b > a)     # every non-empty range starts on a key frame


def test_zstd_decoder_refuses_absurd_declared_sizes(monkeypatch):
    from tezip_b200 import container
    blob = container.zstd_compress(np.zeros(1 << 20, np.uint8))
    assert container.zstd_decompress(blob).size == 1 << 20
    monkeypatch.setenv("TEZIP_MAX_DECODED_BYTES", "1000")
    with pytest.raises(RuntimeError):
        container.zstd_decompress(blob)


def test_run_plan_arrival_ranges_cover_the_sequence_in_order():
    """Streaming loader (compress.run): before each group of windows run_plan asks for the frames that group needs;
    the requests are contiguous, in order, and cover [0, nt) -- checked with a stand-in predictor (no GPU)."""
    import torch

    class FakeNet:
        max_batch, Hp, Wp = 3, 8, 8

        def p0(self, out):
            pass

        def next(self, x, out):
            pass

        def next_chained(self, out):
            pass

    real = codec.ops.pad_normalize
    codec.ops.pad_normalize = lambda frames, idx, Hp, Wp: torch.zeros((len(idx), 8, 8, 3))
    try:
        for nt, p, w, want in ((25, 0, 5, [(0, 15), (15, 25)]), (23, 2, 5, [(0, 17), (17, 23)]), (7, 0, 10, [(0, 7)]),
                               (30, 0, 3, [(0, 9), (9, 18), (18, 27), (27, 30)])):
            calls = []
            plan = codec.plan_from_keys(nt, p, codec.swp_keys(nt, p, w))
            pool = torch.zeros((plan.n_slots, 8, 8, 3))
            codec.run_plan(FakeNet(), torch.zeros((nt, 8, 8, 3), dtype=torch.uint8), plan, pool,
                           lambda a, b: calls.append((a, b)))
            assert calls == want
    finally:
        codec.ops.pad_normalize = real


def test_container_level_override_and_concurrent_frames(tmp_path, monkeypatch):
    """write_container compresses the two files concurrently and honours TEZIP_ZSTD_LEVEL; every level gives single
    frames with content size that read_container (and the reference's zstd.decompress) take."""
    from tezip_b200 import container
    rng = np.random.RandomState(3)
    payload = (rng.geometric(0.3, size=200000) - 1).astype(np.int16)
    key = np.zeros(50000, np.uint8)
    key[:5000] = rng.randint(0, 256, 5000)
    sizes = {}
    for lvl, workers in (("9", 0), ("3", 4), ("1", 2)):
        monkeypatch.setenv("TEZIP_ZSTD_LEVEL", lvl)
        d = str(tmp_path / ("c" + lvl))
        sizes[lvl] = container.write_container(d, ["a.png", "b.png"], True, key, payload, workers=workers)
        names, rgb, k2, p2 = container.read_container(d)
        assert names == ["a.png", "b.png"] and rgb and np.array_equal(k2, key) and np.array_equal(p2, payload)
    monkeypatch.delenv("TEZIP_ZSTD_LEVEL")
    assert container.container_level() == 9
    assert sizes["9"][1] <= sizes["1"][1]
