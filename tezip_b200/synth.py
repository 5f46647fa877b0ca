"""Seeded synthetic frames and random-init PredNet weights.

The same generators feed the oracle, the tests and bench.py, so both sides always see identical inputs.
"""
import numpy as np


def conv_specs(stack_sizes, R_stack_sizes):
    """(key, layer, Cin, Cout) in the reference's weight-list order
    (/root/reference/src/prednet.py:212-227: sorted keys a, ahat, c, f, i, o; layers ascending)."""
    L = len(stack_sizes)
    out = []
    for c in ("a", "ahat", "c", "f", "i", "o"):
        for l in range(L - 1 if c == "a" else L):
            if c == "ahat":
                cin, cout = R_stack_sizes[l], stack_sizes[l]
            elif c == "a":
                cin, cout = 2 * stack_sizes[l], stack_sizes[l + 1]
            else:
                cin = 2 * stack_sizes[l] + R_stack_sizes[l] + (R_stack_sizes[l + 1] if l < L - 1 else 0)
                cout = R_stack_sizes[l]
            out.append((c, l, cin, cout))
    return out


def make_frames(nt, H, W, C=3, seed=1, dtype=np.uint8):
    """f[t,y,x,c] = clip(rint(128 + 80 sin(2pi(x+2t+5c)/32) cos(2pi(y+t+3c)/24) + N(0,2)), 0, 255).

    Never all-zero (the reference decoder cannot tell an all-zero key frame from a non-key frame,
    /root/reference/src/decompress.py:126).  dtype uint16 gives the 16-bit variant
    32768 + 20000*pattern + N(0,64).
    """
    rng = np.random.default_rng(seed)
    t = np.arange(nt, dtype=np.float64)[:, None, None, None]
    y = np.arange(H, dtype=np.float64)[None, :, None, None]
    x = np.arange(W, dtype=np.float64)[None, None, :, None]
    c = np.arange(C, dtype=np.float64)[None, None, None, :]
    pat = np.sin(2 * np.pi * (x + 2 * t + 5 * c) / 32.0) * np.cos(2 * np.pi * (y + t + 3 * c) / 24.0)
    if np.dtype(dtype) == np.uint8:
        f = 128.0 + 80.0 * pat + rng.normal(0.0, 2.0, size=(nt, H, W, C))
        return np.clip(np.rint(f), 0, 255).astype(np.uint8)
    f = 32768.0 + 20000.0 * pat + rng.normal(0.0, 64.0, size=(nt, H, W, C))
    return np.clip(np.rint(f), 0, 65535).astype(np.uint16)


def make_weights(stack_sizes=(3, 48, 96, 192), R_stack_sizes=None, bias="uniform", seed=7):
    """Random-init weights in Keras list order: glorot_uniform kernels U(+-sqrt(6/(9(Cin+Cout)))),
    biases zeros (Keras default, bias='zeros') or U(+-0.1) (bias='uniform')."""
    R_stack_sizes = tuple(R_stack_sizes or stack_sizes)
    rng = np.random.default_rng(seed)
    ws = []
    for (_c, _l, cin, cout) in conv_specs(tuple(stack_sizes), R_stack_sizes):
        lim = np.sqrt(6.0 / (9.0 * (cin + cout)))
        ws.append(rng.uniform(-lim, lim, size=(3, 3, cin, cout)).astype(np.float32))
        if bias == "zeros":
            ws.append(np.zeros((cout,), np.float32))
        else:
            ws.append(rng.uniform(-0.1, 0.1, size=(cout,)).astype(np.float32))
    return ws


def model_json(stack_sizes, R_stack_sizes, Hp, Wp):
    """A prednet_model.json with the two entries compress/decompress read
    (/root/reference/src/compress.py:163,169): layers[0] InputLayer batch_input_shape and
    layers[1] PredNet config (train.py:51-65, prednet.py:310-325)."""
    import json
    return json.dumps({
        "class_name": "Model",
        "config": {"name": "model_1", "layers": [
            {"class_name": "InputLayer", "name": "input_1",
             "config": {"batch_input_shape": [None, 2, Hp, Wp, stack_sizes[0]], "dtype": "float32",
                        "sparse": False, "name": "input_1"}},
            {"class_name": "PredNet", "name": "prednet_1",
             "config": {"name": "prednet_1", "trainable": True, "return_sequences": True,
                        "stack_sizes": list(stack_sizes), "R_stack_sizes": list(R_stack_sizes),
                        "A_filt_sizes": [3] * (len(stack_sizes) - 1), "Ahat_filt_sizes": [3] * len(stack_sizes),
                        "R_filt_sizes": [3] * len(stack_sizes), "pixel_max": 1.0,
                        "error_activation": "relu", "A_activation": "relu", "LSTM_activation": "tanh",
                        "LSTM_inner_activation": "hard_sigmoid", "data_format": "channels_last",
                        "extrap_start_time": None, "output_mode": "error"}}]},
        "keras_version": "2.2.4", "backend": "tensorflow"})


def write_model_dir(path, weights, stack_sizes, R_stack_sizes, Hp, Wp):
    """Model directory: prednet_model.json + prednet_weights.npz (arrays w000..wNNN in Keras list order).
    The reference's prednet_weights.hdf5 needs h5py, which this image lacks."""
    import os
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "prednet_model.json"), "w") as f:
        f.write(model_json(tuple(stack_sizes), tuple(R_stack_sizes), Hp, Wp))
    np.savez(os.path.join(path, "prednet_weights.npz"), **{"w%03d" % i: w for i, w in enumerate(weights)})
