"""The oracle restatement against the golden fixtures recorded from the UNMODIFIED reference
(tests/golden/make_golden.py).  Given the recorded predictions the stream, the key plane and the decoded frames
must be byte-identical; with the live torch-CPU predictor (whose fp32 rounding may differ between CPUs) the
schedule must be identical and the residuals equal except on trunc() boundaries."""
import glob
import os

import numpy as np
import pytest

from oracle import codec_oracle as co
from oracle.prednet_oracle import PredNetOracle
from tezip_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["name"] = os.path.basename(path)[:-4]
    g["stack"] = tuple(int(v) for v in g["stack"])
    for k in ("nt", "H", "W", "seed", "p", "window"):
        g[k] = int(g[k])
    g["window"] = None if g["window"] < 0 else g["window"]
    g["threshold"] = None if float(g["threshold"]) < 0 else float(g["threshold"])
    g["mode"], g["bias"] = str(g["mode"]), str(g["bias"])
    g["bound"] = [float(v) for v in g["bound"]]
    g["entropy"] = bool(g["entropy"])
    g["frames"] = synth.make_frames(g["nt"], g["H"], g["W"], 3, seed=g["seed"])
    return g


def test_fixtures_present():
    assert len(GOLD) >= 8


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_encode_given_recorded_predictions(path):
    g = load(path)
    windows = [(int(f), [g["preds"][int(f) + i] for i in range(int(n))]) for f, n in g["windows"]]
    r = co.encode_windows(g["frames"][np.newaxis], windows, g["p"], g["mode"], g["bound"], g["entropy"])
    assert np.array_equal(r["payload"], g["ref_payload"])
    assert np.array_equal(r["x"], g["x"])
    kp = np.zeros_like(g["frames"])
    kp[g["keys"]] = g["frames"][g["keys"]]
    assert np.array_equal(kp.ravel(), g["ref_key_plane"])


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_decode_given_recorded_predictions(path):
    g = load(path)
    out, info = co.decompress_arrays(g["ref_key_plane"], g["ref_payload"], None, replay_preds=g["preds"])
    assert np.array_equal(out, g["ref_decoded"])
    assert info["keys"] == [int(k) for k in g["keys"]]
    assert info["n_predict_calls"] == int(g["ref_predict_calls"][1])
    if g["bound"][0] == 0:
        assert np.array_equal(out, g["frames"])


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_live_pipeline_against_golden(path):
    g = load(path)
    ws = synth.make_weights(g["stack"], bias=g["bias"], seed=7)
    net = PredNetOracle(ws, g["stack"], g["stack"])
    r = co.compress_arrays(g["frames"], net, g["p"], g["window"], g["threshold"], g["mode"], g["bound"], g["entropy"])
    assert np.abs(r["preds"] - g["preds"]).max() <= 1e-5
    assert r["keys"] == [int(k) for k in g["keys"]]
    assert np.array_equal(r["key_plane"], g["ref_key_plane"])
    assert np.mean(r["x"] == g["x"]) >= 0.999
    if np.array_equal(r["x"], g["x"]):
        assert np.array_equal(r["payload"], g["ref_payload"])
