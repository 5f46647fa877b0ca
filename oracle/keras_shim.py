"""ORACLE (test infrastructure only) -- a numpy stand-in for the slice of Keras 2.2.4 that
/root/reference/src/prednet.py imports, so that the reference's OWN `PredNet.build()`, `get_initial_state()` and
`step()` (prednet.py:143-308) execute unmodified in this container (no TensorFlow / Keras here).

What runs from the reference: the class body of prednet.py -- layer construction and weight-list order
(:192-233), zero states (:143-190), the whole time step (:235-308: concat order, gate wiring, up-sampling, error
units, pooling, clipping).  What this file restates, from the published Keras 2.2.4 / tensorflow_backend source
(third-party code that is NOT under /root/reference): the primitives those lines call --
    K.concatenate / minimum / zeros / zeros_like / sum / dot / reshape / mean / batch_flatten / switch / rnn,
    activations relu / tanh / hard_sigmoid (= clip(0.2*x + 0.5, 0, 1)),
    Conv2D.call  (= activation(bias_add(conv2d(x, kernel, strides 1, 'same'), bias)), HWIO kernels, cross-correlation),
    UpSampling2D.call (nearest, size 2: repeat_elements on rows and columns), MaxPooling2D.call (2x2 / stride 2 / valid),
    legacy `Recurrent`: __init__ keywords, `weights=` applied after build, call() = K.rnn(self.step, inputs,
    self.get_initial_state(inputs)) with return_sequences, get_config().
All arithmetic is float32 numpy (TF computes in floatx = float32); the convolution accumulates tap by tap
(ky, kx ascending) with one float32 matrix product per tap.

Used by oracle/refharness.py (`real_prednet=True`) and tests/golden/make_prednet_golden.py.  Nothing on the GPU
box imports it.
"""
import contextlib
import sys
import types

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------------- backend (K)
class Variable:
    """A mutable tensor holder (K.variable / layer weights)."""

    def __init__(self, value, name=None):
        self.value = np.array(value, dtype=F32)
        self.name = name

    @property
    def shape(self):
        return self.value.shape


def _val(x):
    return x.value if isinstance(x, Variable) else x


def image_data_format():
    return "channels_last"


def backend():
    return "tensorflow"


def zeros_like(x, dtype=None, name=None):
    return np.zeros_like(_val(x), dtype=F32)


def zeros(shape, dtype=None, name=None):
    return np.zeros(tuple(int(s) for s in shape), dtype=F32)


def sum(x, axis=None, keepdims=False):   # noqa: A001  (keras.backend.sum)
    return np.sum(_val(x), axis=axis, keepdims=keepdims, dtype=F32)


def mean(x, axis=None, keepdims=False):
    return np.mean(_val(x), axis=axis, keepdims=keepdims, dtype=F32)


def dot(x, y):
    return np.matmul(_val(x), _val(y)).astype(F32)


def reshape(x, shape):
    return np.reshape(_val(x), tuple(int(s) for s in shape))


def batch_flatten(x):
    x = _val(x)
    return np.reshape(x, (x.shape[0], -1))


def concatenate(tensors, axis=-1):
    return np.concatenate([_val(t) for t in tensors], axis=axis)


def minimum(x, y):
    return np.minimum(_val(x), F32(y) if np.isscalar(y) else _val(y))


def maximum(x, y):
    return np.maximum(_val(x), F32(y) if np.isscalar(y) else _val(y))


def switch(condition, then_expression, else_expression):
    return then_expression if bool(condition) else else_expression


def variable(value, dtype=None, name=None):
    return Variable(value, name)


def int_shape(x):
    return tuple(_val(x).shape)


@contextlib.contextmanager
def name_scope(name):
    yield


def relu(x, alpha=0.0, max_value=None):
    # tensorflow_backend.relu with alpha == 0, max_value None: tf.nn.relu
    return np.maximum(_val(x), F32(0.0))


def tanh(x):
    return np.tanh(_val(x)).astype(F32)


def hard_sigmoid(x):
    # tensorflow_backend.hard_sigmoid: x = (0.2 * x) + 0.5; clip_by_value(x, 0, 1)
    x = (F32(0.2) * _val(x)) + F32(0.5)
    return np.clip(x, F32(0.0), F32(1.0))


def conv2d(x, kernel, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1)):
    """tf.nn.convolution, NHWC, HWIO kernel, cross-correlation (no flip).  'same' with stride 1 and an odd
    kernel pads (k-1)/2 zeros on every side."""
    x, kernel = _val(x), _val(kernel)
    assert tuple(strides) == (1, 1) and tuple(dilation_rate) == (1, 1) and data_format in (None, "channels_last")
    kh, kw, cin, cout = kernel.shape
    assert x.shape[-1] == cin, (x.shape, kernel.shape)
    if padding == "same":
        assert kh % 2 == 1 and kw % 2 == 1
        ph, pw = kh // 2, kw // 2
        xp = np.zeros((x.shape[0], x.shape[1] + 2 * ph, x.shape[2] + 2 * pw, cin), F32)
        xp[:, ph:ph + x.shape[1], pw:pw + x.shape[2]] = x
    else:
        xp = x
    Ho, Wo = xp.shape[1] - kh + 1, xp.shape[2] - kw + 1
    out = np.zeros((x.shape[0], Ho, Wo, cout), F32)
    for ky in range(kh):
        for kx in range(kw):
            out += np.matmul(xp[:, ky:ky + Ho, kx:kx + Wo, :], kernel[ky, kx])   # float32 product per tap
    return out


def bias_add(x, bias, data_format=None):
    return (_val(x) + _val(bias)).astype(F32)


def repeat_elements(x, rep, axis):
    return np.repeat(_val(x), rep, axis=axis)


def resize_images(x, height_factor, width_factor, data_format):
    # tensorflow_backend.resize_images (2.2.4): nearest neighbour == repeat rows, then columns
    assert data_format == "channels_last"
    return repeat_elements(repeat_elements(x, height_factor, axis=1), width_factor, axis=2)


def pool2d(x, pool_size, strides=(1, 1), padding="valid", data_format=None, pool_mode="max"):
    x = _val(x)
    assert pool_mode == "max" and padding == "valid" and tuple(pool_size) == (2, 2) and tuple(strides) == (2, 2)
    B, H, W, C = x.shape
    x = x[:, :H // 2 * 2, :W // 2 * 2]
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def rnn(step_function, inputs, initial_states, go_backwards=False, mask=None, constants=None, unroll=False,
        input_length=None):
    """tensorflow_backend.rnn: iterate axis 1; outputs stacked on axis 1.  -> (last_output, outputs, new_states)."""
    inputs = _val(inputs)
    assert mask is None and not go_backwards
    states = list(initial_states)
    constants = list(constants or [])
    outs = []
    for t in range(inputs.shape[1]):
        out, states = step_function(inputs[:, t], states + constants)
        states = list(states)
        outs.append(out)
    return outs[-1], np.stack(outs, axis=1), states


# ---------------------------------------------------------------------------------------------- activations
def _activation_get(identifier):
    if identifier is None:
        return linear
    if callable(identifier):
        return identifier
    return {"relu": relu, "tanh": tanh, "hard_sigmoid": hard_sigmoid, "linear": linear}[identifier]


def linear(x):
    return x


# ---------------------------------------------------------------------------------------------- layers
class InputSpec:
    def __init__(self, dtype=None, shape=None, ndim=None, max_ndim=None, min_ndim=None, axes=None):
        self.dtype, self.shape, self.ndim = dtype, shape, ndim


class Placeholder:
    """keras.layers.Input(shape=...): only `.shape` is read (compress.py:179)."""

    def __init__(self, shape=None, **kw):
        self.shape = (None,) + tuple(shape)


class _Symbolic:
    def __init__(self, layer, inputs):
        self.layer, self.inputs = layer, inputs


class Layer:
    def __init__(self, weights=None, name=None, trainable=True, **kwargs):
        self._initial_weights = weights
        self.name, self.trainable = name, trainable
        self.built = False
        self.trainable_weights = []

    def add_weight(self, shape, name=None, **kw):
        v = Variable(np.zeros(shape, F32), name)
        self.trainable_weights.append(v)
        return v

    def set_weights(self, weights):
        assert len(weights) == len(self.trainable_weights), (len(weights), len(self.trainable_weights))
        for v, w in zip(self.trainable_weights, weights):
            w = np.asarray(w, dtype=F32)
            assert v.value.shape == w.shape, (v.value.shape, w.shape)   # Layer.set_weights checks shapes
            v.value = w.copy()

    def get_weights(self):
        return [v.value.copy() for v in self.trainable_weights]

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    def __call__(self, inputs, **kw):
        # base Layer.__call__: build on first use, then apply `weights=` (engine/base_layer.py: _initial_weights)
        if not self.built:
            self.build(tuple(inputs.shape))
            self.built = True
            if self._initial_weights is not None:
                self.set_weights(self._initial_weights)
                self._initial_weights = None
        if isinstance(inputs, Placeholder):
            return _Symbolic(self, inputs)
        return self.call(inputs, **kw)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", data_format=None, activation=None,
                 use_bias=True, **kwargs):
        super().__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.strides, self.padding, self.data_format = tuple(strides), padding, data_format or "channels_last"
        self.activation = _activation_get(activation)
        self.use_bias = use_bias

    def build(self, input_shape):
        cin = input_shape[-1] if self.data_format == "channels_last" else input_shape[1]
        self.kernel = self.add_weight(self.kernel_size + (int(cin), self.filters), name="kernel")   # HWIO
        self.bias = self.add_weight((self.filters,), name="bias") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        out = conv2d(inputs, self.kernel, self.strides, self.padding, self.data_format)
        if self.use_bias:
            out = bias_add(out, self.bias, self.data_format)
        return self.activation(out) if self.activation is not None else out


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), data_format=None, **kwargs):
        super().__init__(**kwargs)
        self.size, self.data_format = tuple(size), data_format or "channels_last"

    def call(self, inputs):
        return resize_images(inputs, self.size[0], self.size[1], self.data_format)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", data_format=None, **kwargs):
        super().__init__(**kwargs)
        self.pool_size = tuple(pool_size)
        self.strides = tuple(strides) if strides is not None else self.pool_size
        self.padding, self.data_format = padding, data_format or "channels_last"

    def call(self, inputs):
        return pool2d(inputs, self.pool_size, self.strides, self.padding, self.data_format, "max")


class Recurrent(Layer):
    """keras.legacy.layers.Recurrent (2.2.4), the parts PredNet inherits."""

    def __init__(self, return_sequences=False, return_state=False, go_backwards=False, stateful=False, unroll=False,
                 implementation=0, **kwargs):
        super().__init__(**kwargs)
        self.return_sequences, self.return_state = return_sequences, return_state
        self.go_backwards, self.stateful, self.unroll, self.implementation = go_backwards, stateful, unroll, implementation
        self.supports_masking = True
        self.input_spec = [InputSpec(ndim=3)]
        self.state_spec = None

    def get_constants(self, inputs, training=None):
        return []

    def preprocess_input(self, inputs, training=None):
        return inputs

    def call(self, inputs, mask=None, training=None, initial_state=None):
        assert not self.stateful and initial_state is None
        inputs = np.asarray(inputs, dtype=F32)                  # Model.predict feeds floatx
        initial_state = self.get_initial_state(inputs)
        constants = self.get_constants(inputs, training=None)
        preprocessed = self.preprocess_input(inputs, training=None)
        last_output, outputs, states = rnn(self.step, preprocessed, initial_state, go_backwards=self.go_backwards,
                                           mask=mask, constants=constants, unroll=self.unroll,
                                           input_length=inputs.shape[1])
        return outputs if self.return_sequences else last_output

    def get_config(self):
        config = {"return_sequences": self.return_sequences, "return_state": self.return_state,
                  "go_backwards": self.go_backwards, "stateful": self.stateful, "unroll": self.unroll,
                  "implementation": self.implementation}
        base = super().get_config()
        return dict(list(base.items()) + list(config.items()))


class Model:
    """keras.models.Model(inputs=Input, outputs=layer(Input)) with predict() only."""

    def __init__(self, inputs=None, outputs=None):
        self.input = inputs
        self._layer = outputs.layer
        self.layers = [inputs, outputs.layer]

    def predict(self, x, batch_size=None, verbose=0):
        x = np.asarray(x, dtype=F32)
        bs = int(batch_size) if batch_size else 32
        return np.concatenate([self._layer.call(x[i:i + bs]) for i in range(0, x.shape[0], bs)], axis=0)


def generate_legacy_interface(allowed_positional_args=None, conversions=None, preprocessor=None,
                              value_conversions=None, object_type="class"):
    """keras.legacy.interfaces: the wrapper only renames Keras-1 keyword arguments; with Keras-2 names it calls through."""
    def decorator(func):
        return func
    return decorator


def recurrent_args_preprocessor(args, kwargs):
    return args, kwargs, []


# ---------------------------------------------------------------------------------------------- module injection
_NAMES = ("keras", "keras.backend", "keras.activations", "keras.layers", "keras.engine", "keras.legacy",
          "keras.legacy.interfaces", "keras.models", "keras.preprocessing", "keras.preprocessing.image")


def install():
    """Registers the stand-in `keras` package in sys.modules; returns what was there before (for restore())."""
    saved = {k: sys.modules.get(k) for k in _NAMES}
    me = sys.modules[__name__]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    kb = mod("keras.backend", **{k: getattr(me, k) for k in (
        "image_data_format", "backend", "zeros_like", "zeros", "sum", "mean", "dot", "reshape", "batch_flatten",
        "concatenate", "minimum", "maximum", "switch", "variable", "int_shape", "name_scope", "relu", "tanh",
        "hard_sigmoid", "conv2d", "bias_add", "repeat_elements", "resize_images", "pool2d", "rnn")}, _BACKEND="tensorflow")
    ka = mod("keras.activations", get=_activation_get, relu=relu, tanh=tanh, hard_sigmoid=hard_sigmoid, linear=linear)
    kl = mod("keras.layers", Recurrent=Recurrent, Conv2D=Conv2D, UpSampling2D=UpSampling2D, MaxPooling2D=MaxPooling2D,
             Input=Placeholder, Dense=object, Flatten=object, Layer=Layer)
    ke = mod("keras.engine", InputSpec=InputSpec, Layer=Layer)
    kli = mod("keras.legacy.interfaces", generate_legacy_interface=generate_legacy_interface,
              recurrent_args_preprocessor=recurrent_args_preprocessor)
    kleg = mod("keras.legacy", interfaces=kli)
    km = mod("keras.models", Model=Model)
    kpi = mod("keras.preprocessing.image", Iterator=object)
    kp = mod("keras.preprocessing", image=kpi)
    mod("keras", backend=kb, activations=ka, layers=kl, engine=ke, legacy=kleg, models=km, preprocessing=kp)
    return saved


def restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


REF_PREDNET = "/root/reference/src/prednet.py"


def load_reference_prednet():
    """exec()s the UNMODIFIED /root/reference/src/prednet.py against the stand-in keras; returns its module.
    Call between install() and restore()."""
    src = open(REF_PREDNET, encoding="utf-8").read()
    m = types.ModuleType("ref_prednet")
    m.__file__ = REF_PREDNET
    exec(compile(src, REF_PREDNET, "exec"), m.__dict__)
    return m


class ReferencePredNet:
    """The reference's PredNet class driven the way compress.py:163-173 / decompress.py:75-85 drive it:
    PredNet(weights=..., **layer_config) -> test_prednet(Input) -> Model.predict."""

    def __init__(self, weights, stack_sizes, R_stack_sizes, Hp, Wp, pixel_max=1.0):
        saved = install()
        try:
            self.module = load_reference_prednet()
            L = len(stack_sizes)
            cfg = dict(stack_sizes=tuple(stack_sizes), R_stack_sizes=tuple(R_stack_sizes), A_filt_sizes=(3,) * (L - 1),
                       Ahat_filt_sizes=(3,) * L, R_filt_sizes=(3,) * L, pixel_max=pixel_max, output_mode="prediction",
                       return_sequences=True, data_format="channels_last")        # train.py:51-65 + compress.py:164
            self.layer = self.module.PredNet(weights=list(weights), **cfg)
            inputs = Placeholder(shape=(None, Hp, Wp, stack_sizes[0]))               # compress.py:169-171
            self.model = Model(inputs=inputs, outputs=self.layer(inputs))
        finally:
            restore(saved)

    def predict(self, x, batch_size=None):
        return self.model.predict(x, batch_size)

    def next(self, frames):
        frames = np.asarray(frames, dtype=F32)
        return self.predict(np.stack([frames, np.zeros_like(frames)], axis=1))[:, 1]

    def p0(self, Hp, Wp):
        return self.predict(np.zeros((1, 1, Hp, Wp, self.layer.stack_sizes[0]), F32))[0, 0]
