"""Pins the oracle restatement against the reference ITSELF: runs the unmodified /root/reference/src/compress.py and
decompress.py under oracle/refharness.py (only possible in the build container) and compares byte for byte."""
import tempfile

import numpy as np
import pytest

from oracle import refharness as rh, codec_oracle as co
from oracle.prednet_oracle import PredNetOracle
from tezip_b200 import synth

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference is not mounted on this box")

STACK = (3, 8, 16, 32)


@pytest.mark.parametrize("p,window,thr,mode,bound,entropy", [
    (0, 3, None, "abs", [0.0], True),
    (2, 4, None, "abs", [3.0], True),
    (0, 4, None, "pwrel", [0.05], False),
    (1, None, 0.05, "rel", [0.01], True),
])
def test_restatement_equals_unmodified_reference(p, window, thr, mode, bound, entropy):
    nt, H, W = 9, 12, 20
    Hp, Wp = 16, 24
    ws = synth.make_weights(STACK, bias="uniform", seed=3)
    frames = synth.make_frames(nt, H, W, 3, seed=17)
    net = PredNetOracle(ws, STACK, STACK)
    d = tempfile.mkdtemp()
    synth.write_model_dir(d + "/model", ws, STACK, STACK, Hp, Wp)
    rh.write_png_dir(d + "/img", frames)
    n_c = rh.run_compress(d + "/model", d + "/img", d + "/out", p, window, thr, mode, bound, entropy)
    kp = np.frombuffer(rh.zstd_decompress(open(d + "/out/key_frame.dat", "rb").read()), np.uint8)
    pl = np.frombuffer(rh.zstd_decompress(open(d + "/out/entropy.dat", "rb").read()), np.int16)
    r = co.compress_arrays(frames, net, p, window, thr, mode, bound, entropy)
    assert np.array_equal(kp, r["key_plane"])
    assert np.array_equal(pl, r["payload"])
    assert n_c == nt - 1                                       # compress makes nt-1 predict calls
    n_d = rh.run_decompress(d + "/model", d + "/out", d + "/dec")
    names = [l.strip() for l in open(d + "/out/filename.txt")][1:]
    dec = rh.read_png_dir(d + "/dec", names)
    out, info = co.decompress_arrays(r["key_plane"], r["payload"], net)
    assert np.array_equal(dec, out)
    assert info["n_predict_calls"] == n_d


@pytest.mark.parametrize("p,window,thr,mode,bound", [(0, 3, None, "abs", [2.0]), (1, None, 0.05, "abs", [0.0])])
def test_whole_reference_with_its_own_prednet(p, window, thr, mode, bound):
    """compress.py + decompress.py + prednet.py, all three unmodified (the PredNet class runs over the numpy Keras
    stand-in, oracle/keras_shim.py): the codec restatement fed by the same class reproduces the files byte for byte,
    and oracle/prednet_oracle.py agrees with the reference's predictions to 1e-6."""
    from oracle.keras_shim import ReferencePredNet
    nt, H, W = 8, 12, 20
    Hp, Wp = 16, 24
    ws = synth.make_weights(STACK, bias="uniform", seed=3)
    frames = synth.make_frames(nt, H, W, 3, seed=19)
    d = tempfile.mkdtemp()
    synth.write_model_dir(d + "/model", ws, STACK, STACK, Hp, Wp)
    rh.write_png_dir(d + "/img", frames)
    rh._State.predict_log = log = []
    try:
        rh.run_compress(d + "/model", d + "/img", d + "/out", p, window, thr, mode, bound, True, real_prednet=True)
    finally:
        rh._State.predict_log = None
    kp = np.frombuffer(rh.zstd_decompress(open(d + "/out/key_frame.dat", "rb").read()), np.uint8)
    pl = np.frombuffer(rh.zstd_decompress(open(d + "/out/entropy.dat", "rb").read()), np.int16)
    net = ReferencePredNet(ws, STACK, STACK, Hp, Wp)
    r = co.compress_arrays(frames, net, p, window, thr, mode, bound, True)
    assert np.array_equal(kp, r["key_plane"]) and np.array_equal(pl, r["payload"])
    rh.run_decompress(d + "/model", d + "/out", d + "/dec", real_prednet=True)
    names = [l.strip() for l in open(d + "/out/filename.txt")][1:]
    out, _info = co.decompress_arrays(r["key_plane"], r["payload"], net)
    assert np.array_equal(rh.read_png_dir(d + "/dec", names), out)
    onet = PredNetOracle(ws, STACK, STACK)
    assert len(log) == nt - 1
    for x, y in log:                                            # every predict call the reference made
        assert np.abs(onet.predict(x) - y).max() <= 1e-6


def test_prednet_properties():
    """X_hat[0,0] is input independent; X_hat[0,1] ignores the t=1 input; no state crosses calls."""
    ws = synth.make_weights(STACK, bias="uniform", seed=3)
    net = PredNetOracle(ws, STACK, STACK)
    rng = np.random.default_rng(0)
    a, b = rng.random((1, 2, 16, 24, 3), dtype=np.float32), rng.random((1, 2, 16, 24, 3), dtype=np.float32)
    pa, pb = net.predict(a), net.predict(b)
    assert np.array_equal(pa[0, 0], pb[0, 0])
    b2 = b.copy(); b2[0, 0] = a[0, 0]
    assert np.array_equal(net.predict(b2)[0, 1], pa[0, 1])
    assert np.array_equal(net.predict(a), pa)
    assert sum(w.size for w in synth.make_weights((3, 48, 96, 192))) == 6915948
