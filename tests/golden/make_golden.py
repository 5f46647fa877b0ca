"""Generates tests/golden/*.npz by running the UNMODIFIED reference -- compress.py, decompress.py AND prednet.py
(oracle/refharness.py with real_prednet=True: the reference's own PredNet class executes over the numpy Keras
stand-in of oracle/keras_shim.py) -- in the build container, and checks at generation time that the restatement
in oracle/codec_oracle.py reproduces the reference's files byte for byte.

    python tests/golden/make_golden.py            (needs /root/reference; rewrites the fixtures)

Each fixture holds the inputs' recipe (seeds, shapes, flags), and from the REFERENCE run: the key plane, the
entropy.dat payload before zstd, the decoded frames; plus the per-frame predictions the reference used
(float32, what the reference's prednet.py returned) so that the GPU codec kernels can be tested against the
reference's stream given identical predictions (tests/test_gpu_golden.py).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refharness as rh, codec_oracle as co          # noqa: E402
from oracle.keras_shim import ReferencePredNet                    # noqa: E402
from tezip_b200 import synth                                      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STACK = (3, 16, 32, 64)
FULL = (3, 48, 96, 192)          # train.py:51

CASES = {
    # name: (nt, H, W, seed, bias, p, window, threshold, mode, bound, entropy)
    "lossless_w5": (12, 20, 28, 1, "uniform", 0, 5, None, "abs", [0.0], True),
    "warmup2_abs2": (12, 20, 28, 2, "uniform", 2, 5, None, "abs", [2.0], True),
    "lastkey_w5": (11, 16, 24, 3, "uniform", 0, 5, None, "abs", [0.0], True),
    "abs2p55": (10, 16, 24, 4, "zeros", 0, 4, None, "abs", [2.55], True),
    "rel": (10, 20, 28, 5, "uniform", 0, 4, None, "rel", [0.02], True),
    "absrel_noentropy": (12, 20, 28, 6, "uniform", 1, 4, None, "absrel", [3.0, 0.1], False),
    "pwrel": (10, 16, 24, 7, "uniform", 0, 5, None, "pwrel", [0.03], True),
    "dwp": (12, 16, 24, 8, "uniform", 0, None, "calibrate", "abs", [0.0], True),
    # the BASELINE frame shape with the train.py:51 architecture (a 12-tuple: the last entry is the stack)
    "full_128x160_abs2": (7, 128, 160, 9, "uniform", 0, 3, None, "abs", [2.0], True, FULL),
}


def calibrate_threshold(frames, net, Hp, Wp):
    """from the cumulative window MSE of an unbounded window; halfway between its values at the
    third and fourth step, so that the close decision (compress.py:249) does not sit on a float32 rounding edge."""
    r = co.compress_arrays(frames, net, 0, None, 1e9, "abs", [0.0], True)
    nt, H, W, _ = frames.shape
    pads = np.zeros((nt, Hp, Wp, 3))
    pads[:, :H, :W] = frames.astype(np.float32) / 255
    m3 = float(np.mean((pads[1:4] - r["preds"][1:4].astype(np.float64)) ** 2))
    m4 = float(np.mean((pads[1:5] - r["preds"][1:5].astype(np.float64)) ** 2))
    assert abs(m4 - m3) > 1e-4
    return 0.5 * (m3 + m4)


def main():
    assert rh.available(), "/root/reference is not mounted"
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        nt, H, W, seed, bias, p, window, thr, mode, bound, entropy = case[:11]
        STACK = case[11] if len(case) > 11 else globals()["STACK"]
        Hp, Wp = co.padding_size(H), co.padding_size(W)
        ws = synth.make_weights(STACK, bias=bias, seed=7)
        frames = synth.make_frames(nt, H, W, 3, seed=seed)
        net = ReferencePredNet(ws, STACK, STACK, Hp, Wp)          # the reference's own prednet.py
        if thr == "calibrate":
            thr = calibrate_threshold(frames, net, Hp, Wp)
        d = tempfile.mkdtemp()
        synth.write_model_dir(d + "/model", ws, STACK, STACK, Hp, Wp)
        rh.write_png_dir(d + "/img", frames)
        n_c = rh.run_compress(d + "/model", d + "/img", d + "/out", p, window, thr, mode, bound, entropy,
                              real_prednet=True)
        kp = np.frombuffer(rh.zstd_decompress(open(d + "/out/key_frame.dat", "rb").read()), np.uint8)
        pl = np.frombuffer(rh.zstd_decompress(open(d + "/out/entropy.dat", "rb").read()), np.int16)
        n_d = rh.run_decompress(d + "/model", d + "/out", d + "/dec", real_prednet=True)
        names = [l.strip() for l in open(d + "/out/filename.txt")][1:]
        dec = rh.read_png_dir(d + "/dec", names)
        r = co.compress_arrays(frames, net, p, window, thr, mode, bound, entropy)
        out, info = co.decompress_arrays(r["key_plane"], r["payload"], net)
        assert np.array_equal(kp, r["key_plane"]) and np.array_equal(pl, r["payload"]) and np.array_equal(dec, out), name
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            stack=np.array(STACK), nt=nt, H=H, W=W, seed=seed, bias=bias, p=p,
            window=-1 if window is None else window, threshold=-1.0 if thr is None else thr, mode=mode,
            bound=np.array(bound, np.float64), entropy=entropy,
            ref_key_plane=kp, ref_payload=pl, ref_decoded=dec, ref_predict_calls=np.array([n_c, n_d]),
            preds=r["preds"].astype(np.float32), keys=np.array(r["keys"]), windows=np.array(r["windows"]),
            x=r["x"])
        print(name, "keys", r["keys"], "calls", n_c, n_d, "maxerr", int(np.abs(dec.astype(int) - frames).max()))


if __name__ == "__main__":
    main()
