"""Host-side mirror of the reference's predictor interface, backed by the CUDA library.

`PredNet(stack_sizes, R_stack_sizes, A_filt_sizes, Ahat_filt_sizes, R_filt_sizes, ..., weights=...)` takes the
constructor arguments of /root/reference/src/prednet.py:77-82 and the `weights=` list the reference passes at
compress.py:168 / decompress.py:80; `.predict(x)` has keras `Model.predict` semantics for the only protocol
the drivers use (compress.py:191-197,224-229; decompress.py:141-143,150-178).  The hot-path entry points are
`next()` (= predict([A, 0])[:, 1]) and `p0()` on device tensors.
"""
import ctypes
import json
import os

import numpy as np
import torch

from . import _lib
from .synth import conv_specs


def load_model_dir(weights_dir):
    """prednet_model.json + weights, as compress.py:143-173 reads them.  Returns (layer_config, (Hp, Wp, C),
    weights list).  Weights come from prednet_weights.npz (arrays in Keras list order); the reference's
    prednet_weights.hdf5 is read only if h5py is importable (it is not in this image)."""
    json_file = os.path.join(weights_dir, "prednet_model.json")
    with open(json_file, "r") as f:
        js = json.load(f)
    layers = js["config"]["layers"] if isinstance(js["config"], dict) else js["config"]
    bis = layers[0]["config"]["batch_input_shape"]                       # compress.py:169
    cfg = dict(layers[1]["config"])                                      # compress.py:163
    npz = os.path.join(weights_dir, "prednet_weights.npz")
    h5 = os.path.join(weights_dir, "prednet_weights.hdf5")
    if os.path.exists(npz):
        z = np.load(npz)
        weights = [z[k] for k in sorted(z.files)]
    elif os.path.exists(h5):
        weights = _load_hdf5_weights(h5, cfg)
    else:
        raise OSError(h5)
    return cfg, (int(bis[2]), int(bis[3]), int(bis[4])), weights


def _load_hdf5_weights(path, cfg):
    try:
        import h5py
    except ImportError:
        raise _lib.TezipError("reading %s needs h5py, which is not installed; convert it once with "
                              "scripts/convert_keras_weights.py where h5py exists" % path)
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f           # train.py:109 saves a full-model file
        layer = [k for k in g.keys() if "prednet" in k.lower()][0]
        names = [n.decode() if isinstance(n, bytes) else n for n in g[layer].attrs["weight_names"]]
        return [np.asarray(g[layer][n]) for n in names]


class PredNet:
    def __init__(self, stack_sizes, R_stack_sizes, A_filt_sizes=None, Ahat_filt_sizes=None, R_filt_sizes=None,
                 pixel_max=1., error_activation='relu', A_activation='relu', LSTM_activation='tanh',
                 LSTM_inner_activation='hard_sigmoid', output_mode='prediction', extrap_start_time=None,
                 data_format='channels_last', weights=None, input_hw=None, max_batch=128, device=0,
                 fp32_direct=False, **kwargs):
        L = len(stack_sizes)
        if len(R_stack_sizes) != L:
            raise ValueError('len(R_stack_sizes) must equal len(stack_sizes)')          # prednet.py:85
        for name, fs, n in (("A_filt_sizes", A_filt_sizes, L - 1), ("Ahat_filt_sizes", Ahat_filt_sizes, L),
                            ("R_filt_sizes", R_filt_sizes, L)):
            if fs is not None and (len(fs) != n or any(int(k) != 3 for k in fs)):
                raise ValueError("%s: only 3x3 filters are supported (train.py:53-55)" % name)
        if (error_activation, A_activation, LSTM_activation, LSTM_inner_activation) != \
                ('relu', 'relu', 'tanh', 'hard_sigmoid'):
            raise ValueError("only relu/relu/tanh/hard_sigmoid activations are supported (prednet.py:79-80)")
        if output_mode not in ('prediction', 'error') or extrap_start_time is not None or \
                data_format != 'channels_last':
            raise ValueError("only output_mode='prediction', channels_last, no extrapolation (compress.py:164)")
        if weights is None or input_hw is None:
            raise ValueError("weights= and input_hw=(Hp, Wp) are required")
        self.stack_sizes, self.R_stack_sizes = tuple(map(int, stack_sizes)), tuple(map(int, R_stack_sizes))
        self.pixel_max = float(pixel_max)
        self.Hp, self.Wp = int(input_hw[0]), int(input_hw[1])
        self.C = self.stack_sizes[0]
        self.max_batch = int(max_batch)
        self.device = torch.device("cuda", int(device))
        specs = conv_specs(self.stack_sizes, self.R_stack_sizes)
        if len(weights) != 2 * len(specs):
            raise ValueError("expected %d weight arrays, got %d" % (2 * len(specs), len(weights)))
        ws = []
        for n, (_c, _l, cin, cout) in enumerate(specs):
            k = np.ascontiguousarray(weights[2 * n], dtype=np.float32)
            b = np.ascontiguousarray(weights[2 * n + 1], dtype=np.float32)
            if k.shape != (3, 3, cin, cout) or b.shape != (cout,):
                raise ValueError("weight %d has shape %s, expected %s" % (2 * n, k.shape, (3, 3, cin, cout)))
            ws += [k, b]
        lib = _lib.load()
        cfg = _lib.PrednetConfig()
        cfg.n_layers = L
        for i in range(L):
            cfg.stack_sizes[i], cfg.r_stack_sizes[i] = self.stack_sizes[i], self.R_stack_sizes[i]
        cfg.Hp, cfg.Wp, cfg.pixel_max = self.Hp, self.Wp, self.pixel_max
        cfg.max_batch, cfg.device = self.max_batch, int(device)
        cfg.flags = _lib.TZ_PREDNET_FP32_DIRECT if fp32_direct else 0
        ptrs = (ctypes.c_void_p * len(ws))(*[w.ctypes.data for w in ws])
        sizes = (ctypes.c_longlong * len(ws))(*[w.size for w in ws])
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.tz_prednet_create(ctypes.byref(cfg), ptrs, sizes, len(ws), ctypes.byref(h)),
                       "tz_prednet_create")
        self._h, self._lib = h, lib

    @classmethod
    def from_model_dir(cls, weights_dir, **kw):
        cfg, (Hp, Wp, _C), weights = load_model_dir(weights_dir)
        keep = {k: cfg[k] for k in ("stack_sizes", "R_stack_sizes", "A_filt_sizes", "Ahat_filt_sizes", "R_filt_sizes",
                                    "pixel_max", "error_activation", "A_activation", "LSTM_activation",
                                    "LSTM_inner_activation") if k in cfg}
        return cls(weights=weights, input_hw=(Hp, Wp), **keep, **kw)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tz_prednet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ hot-path entry points (device tensors)
    def frame_shape(self):
        return (self.Hp, self.Wp, self.C)

    def p0(self, out=None):
        """P0 = predict(anything)[0,0] as a device tensor [Hp,Wp,C] f32."""
        if out is None:
            out = torch.empty(self.frame_shape(), dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.tz_prednet_p0(self._h, _lib.ptr(out), ctypes.c_void_p(st)), "tz_prednet_p0")
        return out

    def next(self, x, out=None):
        """x f32 [B,Hp,Wp,C] (device, contiguous) -> predict([x, 0])[:, 1] as a device tensor."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape[1:]) == self.frame_shape()
        B = x.shape[0]
        if out is None:
            out = torch.empty_like(x)
        assert out.is_contiguous() and out.shape == x.shape and out.data_ptr() != x.data_ptr()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.tz_prednet_next(self._h, _lib.ptr(x), _lib.ptr(out), B, ctypes.c_void_p(st)),
                   "tz_prednet_next")
        return out

    def next_chained(self, out):
        """next() of the first out.shape[0] frames of the prediction the previous next()/next_chained() call on
        this net wrote (compress.py:222-229 feeds X_hat[0, 1] straight back); the caller must not have modified
        that tensor.  Bit-identical to next(previous_out[:B], out); saves one kernel per step."""
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
        assert tuple(out.shape[1:]) == self.frame_shape()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.tz_prednet_next_chained(self._h, _lib.ptr(out), out.shape[0], ctypes.c_void_p(st)),
                   "tz_prednet_next_chained")
        return out

    def kernels(self):
        """[(name, algorithmic FLOPs per frame)] of the launches inside one next()."""
        out = []
        for i in range(self._lib.tz_prednet_kernel_count(self._h)):
            name = ctypes.create_string_buffer(64)
            fl = ctypes.c_double()
            _lib.check(self._lib.tz_prednet_kernel_info(self._h, i, name, 64, ctypes.byref(fl)), "tz_prednet_kernel_info")
            out.append((name.value.decode(), fl.value))
        return out

    def next_timed(self, x, out):
        """One next() with CUDA events between its launches -> per-kernel device ms (synchronous)."""
        n = self._lib.tz_prednet_kernel_count(self._h)
        ms = (ctypes.c_float * n)()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.tz_prednet_next_timed(self._h, _lib.ptr(x), _lib.ptr(out), x.shape[0], ctypes.c_void_p(st),
                                                   ms, n), "tz_prednet_next_timed")
        return list(ms)

    def flops_per_frame(self):
        return float(self._lib.tz_prednet_flops_per_frame(self._h))

    def device_bytes(self):
        return int(self._lib.tz_prednet_device_bytes(self._h))

    # ------------------------------------------------------------------ keras-compatible call (host arrays)
    def predict(self, x, batch_size=None):
        """keras Model.predict for the drivers' protocol: x [B,T,Hp,Wp,C] with T=1 (-> P0) or T=2 with a zero
        second frame (-> [P0, next(x[:,0])]).  Returns float32 numpy like keras."""
        x = np.asarray(x)
        if x.ndim != 5 or x.shape[1] not in (1, 2) or tuple(x.shape[2:]) != self.frame_shape():
            raise ValueError("predict expects [B,1|2,%d,%d,%d]" % self.frame_shape())
        B, T = x.shape[0], x.shape[1]
        out = np.empty((B, T) + self.frame_shape(), np.float32)
        out[:, 0] = self.p0().cpu().numpy()[None]
        if T == 2:
            if np.any(x[:, 1] != 0):
                raise ValueError("only predict([frame, zeros]) is supported (compress.py:225-226)")
            a = torch.from_numpy(np.ascontiguousarray(x[:, 0], dtype=np.float32)).to(self.device)
            res = []
            for b0 in range(0, B, self.max_batch):
                res.append(self.next(a[b0:b0 + self.max_batch].contiguous()).cpu())
            out[:, 1] = torch.cat(res).numpy()
        return out
