"""Convert the reference's Keras weight file to the layout tezip_b200 loads.

    python scripts/convert_keras_weights.py MODEL_DIR            # reads  MODEL_DIR/prednet_weights.hdf5
                                                                 # writes MODEL_DIR/prednet_weights.npz

Run it where h5py exists (the machine that trained the model); the GPU box needs neither h5py nor Keras:
`prednet_model.json` is parsed as plain JSON and the .npz holds the 46 arrays of
`train_model.layers[1].get_weights()` (/root/reference/src/compress.py:157-168) as w000 ... w045, i.e. for
c in (a, ahat, c, f, i, o), for l ascending: kernel [3,3,Cin,Cout], bias [Cout]
(/root/reference/src/prednet.py:212-227).

Keras 2.2.4 HDF5 layout (keras/engine/saving.py, save_weights_to_hdf5_group): the root group -- or the
`model_weights` group of a full-model file written by model.save (/root/reference/src/train.py:109) -- has an attribute
`layer_names`; each layer is a group with an attribute `weight_names` listing its datasets IN get_weights() ORDER.
The PredNet layer is the only layer of the model that owns weights.
"""
import os
import sys

import numpy as np


def _names(attr):
    return [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in attr]


def prednet_weights(root):
    """root: an h5py File/Group (or anything with .attrs / [] of the same shape) -> list of numpy arrays."""
    g = root["model_weights"] if "layer_names" not in root.attrs and "model_weights" in root else root
    if "layer_names" not in g.attrs:
        raise ValueError("not a Keras weight file: no layer_names attribute")
    owners = []
    for lname in _names(g.attrs["layer_names"]):
        wn = _names(g[lname].attrs.get("weight_names", []))
        if wn:
            owners.append((lname, wn))
    if len(owners) != 1:
        raise ValueError("expected exactly one layer with weights (the PredNet layer), found %r"
                         % [o[0] for o in owners])
    lname, wn = owners[0]
    ws = [np.asarray(g[lname][n], dtype=np.float32) for n in wn]
    if len(ws) % 2 or any(ws[i].ndim != 4 or ws[i + 1].ndim != 1 or ws[i].shape[3] != ws[i + 1].shape[0]
                          for i in range(0, len(ws), 2)):
        raise ValueError("layer %s does not look like PredNet: expected (kernel, bias) pairs" % lname)
    return ws


def convert(model_dir, src="prednet_weights.hdf5", dst="prednet_weights.npz"):
    import h5py
    with h5py.File(os.path.join(model_dir, src), "r") as f:
        ws = prednet_weights(f)
    np.savez(os.path.join(model_dir, dst), **{"w%03d" % i: w for i, w in enumerate(ws)})
    return len(ws)


if __name__ == "__main__":
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    print("wrote %d arrays" % convert(sys.argv[1]))
