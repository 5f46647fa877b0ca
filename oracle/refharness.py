"""ORACLE (test infrastructure only) -- executes the UNMODIFIED reference drivers under stub modules.

Only usable where /root/reference is mounted (the build container).  It is what pins the restatement in
codec_oracle.py and what generates tests/golden/*.npz (tests/golden/make_golden.py).  Nothing that runs
on the GPU box (-m gpu tests, smoke(), bench.py) may depend on it.

Recipe: fake modules keras{,.backend,.models,.layers,.preprocessing.image}, prednet,
hickle, numba and zstd (libzstd.so.1 through ctypes) are injected into sys.modules, then
/root/reference/src/{data_utils,compress,decompress}.py are exec'd with ONE text patch,
`.tostring()` -> `.tobytes()` (NumPy >= 2.3 removed ndarray.tostring; used at compress.py:273,395).
Model.predict is backed by oracle/prednet_oracle.py -- or, with real_prednet=True, by the reference's OWN
/root/reference/src/prednet.py executed unmodified over the numpy Keras stand-in of oracle/keras_shim.py (then
`from prednet import PredNet`, `Model`, `Input` inside compress.py / decompress.py are the real class and the
stand-in's eager Model).  GPU_FLAG is always False.
"""
import ctypes
import io
import contextlib
import json
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"


def available():
    return os.path.isfile(os.path.join(REF_SRC, "compress.py"))


# ---------------------------------------------------------------------------------------------- zstd stub
_z = None


def _zlib():
    global _z
    if _z is None:
        _z = ctypes.CDLL("libzstd.so.1")
        _z.ZSTD_compressBound.restype = ctypes.c_size_t
        _z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        _z.ZSTD_compress.restype = ctypes.c_size_t
        _z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        _z.ZSTD_decompress.restype = ctypes.c_size_t
        _z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        _z.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        _z.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _z.ZSTD_isError.restype = ctypes.c_uint
        _z.ZSTD_isError.argtypes = [ctypes.c_size_t]
    return _z


def zstd_compress(data, level=3):
    """python-zstd 1.4.5.1 `zstd.compress(data, level)`: one frame carrying its content size."""
    z = _zlib()
    data = bytes(data)
    cap = z.ZSTD_compressBound(len(data))
    dst = ctypes.create_string_buffer(cap)
    n = z.ZSTD_compress(dst, cap, data, len(data), int(level))
    if z.ZSTD_isError(n):
        raise RuntimeError("ZSTD_compress failed")
    return dst.raw[:n]


def zstd_decompress(data):
    z = _zlib()
    data = bytes(data)
    size = z.ZSTD_getFrameContentSize(data, len(data))
    if size >= (1 << 62):
        raise RuntimeError("zstd frame without content size")
    dst = ctypes.create_string_buffer(max(int(size), 1))
    n = z.ZSTD_decompress(dst, int(size), data, len(data))
    if z.ZSTD_isError(n):
        raise RuntimeError("ZSTD_decompress failed")
    return dst.raw[:n]


# ---------------------------------------------------------------------------------------------- keras stubs
class _State:
    predictor = None       # object with .predict(x, batch_size)
    weights = None         # list of arrays, Keras order
    Hp = Wp = C = None
    n_predict_calls = 0
    predict_log = None     # optional list receiving (input copy, output copy)


class _Shape(tuple):
    pass


class _FakeTrainLayer:
    def __init__(self, cfg=None, bis=None):
        self._cfg, self.batch_input_shape = cfg, bis

    def get_config(self):
        return dict(self._cfg)

    def get_weights(self):
        return _State.weights


class _FakeTrainModel:
    def __init__(self, js):
        layers = json.loads(js)["config"]["layers"]
        self.layers = [_FakeTrainLayer(bis=tuple(layers[0]["config"]["batch_input_shape"])),
                       _FakeTrainLayer(cfg=layers[1]["config"])]

    def load_weights(self, path):
        # the stub keeps weights in an .npz beside the (absent) .hdf5: no HDF5 library in this image
        npz = os.path.join(os.path.dirname(path), "prednet_weights.npz")
        if not os.path.exists(npz):
            raise OSError(path)
        z = np.load(npz)
        _State.weights = [z[k] for k in sorted(z.files)]


class _FakePredNet:
    def __init__(self, *a, weights=None, **cfg):
        self.cfg = cfg
        self.weights = weights

    def __call__(self, inputs):
        return ("predictions", self, inputs)


class _FakeInput:
    def __init__(self, shape):
        self.shape = _Shape((None,) + tuple(shape))


class _FakeTestModel:
    def __init__(self, inputs=None, outputs=None):
        self.input = inputs
        _tag, pn, _ = outputs
        from oracle.prednet_oracle import PredNetOracle
        c = pn.cfg
        self._net = PredNetOracle(pn.weights, c["stack_sizes"], c["R_stack_sizes"], c.get("pixel_max", 1.0))
        if _State.predictor is not None:
            self._net = _State.predictor

    def predict(self, x, batch_size=None):
        _State.n_predict_calls += 1
        out = self._net.predict(x, batch_size)
        if _State.predict_log is not None:
            _State.predict_log.append((np.array(x, copy=True), np.array(out, copy=True)))
        return out


def _install_stubs(real_prednet=False):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in
             ("keras", "keras.backend", "keras.models", "keras.layers", "keras.preprocessing",
              "keras.preprocessing.image", "prednet", "hickle", "numba", "zstd", "data_utils")}
    kb = mod("keras.backend", image_data_format=lambda: "channels_last")
    km = mod("keras.models", Model=_FakeTestModel,
             model_from_json=lambda js, custom_objects=None: _FakeTrainModel(js))
    kl = mod("keras.layers", Input=lambda shape=None: _FakeInput(shape), Dense=object, Flatten=object)
    kpi = mod("keras.preprocessing.image", Iterator=object)
    kp = mod("keras.preprocessing", image=kpi)
    mod("keras", backend=kb, models=km, layers=kl, preprocessing=kp)
    mod("prednet", PredNet=_FakePredNet)
    if real_prednet:
        # the reference's own prednet.py over the numpy Keras stand-in; model_from_json stays the JSON/npz reader
        from oracle import keras_shim
        saved.update({k: v for k, v in keras_shim.install().items() if k not in saved})
        sys.modules["keras.models"].model_from_json = lambda js, custom_objects=None: _FakeTrainModel(js)

        class _CountingModel(keras_shim.Model):
            def predict(self, x, batch_size=None, verbose=0):
                _State.n_predict_calls += 1
                out = super().predict(x, batch_size)
                if _State.predict_log is not None:
                    _State.predict_log.append((np.array(x, copy=True), np.array(out, copy=True)))
                return out

        sys.modules["keras.models"].Model = _CountingModel
        sys.modules["prednet"] = keras_shim.load_reference_prednet()
    mod("hickle", load=lambda *a, **k: None)
    mod("numba", cuda=types.SimpleNamespace(select_device=lambda i: None, close=lambda: None))
    mod("zstd", compress=zstd_compress, decompress=zstd_decompress)
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _load_ref(name):
    path = os.path.join(REF_SRC, name + ".py")
    src = open(path, encoding="utf-8").read().replace(".tostring()", ".tobytes()")
    m = types.ModuleType("ref_" + name)
    m.__file__ = path
    exec(compile(src, path, "exec"), m.__dict__)
    return m


class RefModules:
    """Context manager giving the reference's own compress / decompress / data_utils modules."""

    def __init__(self, real_prednet=False):
        self.real_prednet = real_prednet

    def __enter__(self):
        if not available():
            raise RuntimeError("/root/reference is not mounted here")
        self._saved = _install_stubs(self.real_prednet)
        self.data_utils = _load_ref("data_utils")
        sys.modules["data_utils"] = self.data_utils
        self.compress = _load_ref("compress")
        self.decompress = _load_ref("decompress")
        return self

    def __exit__(self, *exc):
        _restore(self._saved)
        return False


def run_compress(model_dir, img_dir, out_dir, p, window, threshold, mode, bound, entropy=True, verbose=False,
                 predictor=None, real_prednet=False):
    """compress.run exactly as tezip.py:54/56 calls it (GPU_FLAG False). Returns number of predict calls."""
    _State.predictor, _State.n_predict_calls = predictor, 0
    with RefModules(real_prednet) as ref, contextlib.redirect_stdout(io.StringIO() if not verbose else sys.stdout):
        ref.compress.run(model_dir, img_dir, out_dir, p, window, threshold, mode, list(bound), False, verbose,
                         entropy)
    return _State.n_predict_calls


def run_decompress(model_dir, comp_dir, out_dir, verbose=False, predictor=None, real_prednet=False):
    _State.predictor, _State.n_predict_calls = predictor, 0
    with RefModules(real_prednet) as ref, contextlib.redirect_stdout(io.StringIO() if not verbose else sys.stdout):
        ref.decompress.run(model_dir, comp_dir, out_dir, False, verbose)
    return _State.n_predict_calls


def write_png_dir(path, frames):
    """frames u8 [nt,H,W,3] -> %05d.png files (sorted() order == frame order)."""
    from PIL import Image
    os.makedirs(path, exist_ok=True)
    for i, f in enumerate(frames):
        Image.fromarray(np.ascontiguousarray(f)).save(os.path.join(path, "%05d.png" % i))


def read_png_dir(path, names):
    from PIL import Image
    return np.stack([np.array(Image.open(os.path.join(path, n))) for n in names])
