"""oracle/prednet_oracle.py against predictions of the reference's OWN prednet.py (tests/golden/prednet/*.npz, made by
tests/golden/make_prednet_golden.py: /root/reference/src/prednet.py:143-308 executed unmodified over the numpy Keras
stand-in).  Tolerance 1e-6 on predictions in [0, 1] (both are float32; only the summation order inside a
convolution differs).  Where /root/reference is mounted the fixtures themselves are re-derived and compared."""
import glob
import os

import numpy as np
import pytest

from oracle.prednet_oracle import PredNetOracle
from tezip_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prednet", "*.npz")))
TOL = 1e-6


def load(path):
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["stack"] = tuple(int(v) for v in g["stack"])
    for k in ("Hp", "Wp", "B", "seed", "wseed"):
        g[k] = int(g[k])
    g["bias"] = str(g["bias"])
    g["weights"] = synth.make_weights(g["stack"], bias=g["bias"], seed=g["wseed"])
    g["frames"] = synth.make_frames(g["B"], g["Hp"], g["Wp"], g["stack"][0], seed=g["seed"]).astype(np.float32) / 255
    return g


def test_fixtures_present():
    assert len(GOLD) >= 5


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_equals_reference_prednet(path):
    g = load(path)
    net = PredNetOracle(g["weights"], g["stack"], g["stack"])
    assert np.abs(net.p0(g["Hp"], g["Wp"]) - g["p0"]).max() <= TOL
    n1 = net.next(g["frames"])
    assert np.abs(n1 - g["next1"]).max() <= TOL
    assert np.abs(net.next(g["next1"]) - g["next2"]).max() <= TOL      # the fed-back step, from the reference's own input
    assert np.abs(net.next(n1) - g["next2"]).max() <= 2 * TOL


@pytest.mark.parametrize("name", ["tiny_uniform", "mono"])
def test_fixture_regenerates_from_reference_source(name):
    from oracle import keras_shim
    if not os.path.isfile(keras_shim.REF_PREDNET):
        pytest.skip("/root/reference is not mounted on this box")
    g = load(os.path.join(os.path.dirname(__file__), "golden", "prednet", name + ".npz"))
    ref = keras_shim.ReferencePredNet(g["weights"], g["stack"], g["stack"], g["Hp"], g["Wp"])
    # the class that ran is the reference's: its weight list has the order of prednet.py:212-227
    specs = synth.conv_specs(g["stack"], g["stack"])
    shapes = [tuple(v.value.shape) for v in ref.layer.trainable_weights]
    assert shapes == [s for (_c, _l, cin, cout) in specs for s in ((3, 3, cin, cout), (cout,))]
    assert np.abs(ref.next(g["frames"]) - g["next1"]).max() <= 1e-7
    assert np.abs(ref.p0(g["Hp"], g["Wp"]) - g["p0"]).max() <= 1e-7
