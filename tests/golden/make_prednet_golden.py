"""Generates tests/golden/prednet/*.npz: predictions of the reference's OWN /root/reference/src/prednet.py
(`PredNet.build()` / `get_initial_state()` / `step()`, :143-308, executed unmodified over the numpy Keras stand-in of
oracle/keras_shim.py) driven exactly as compress.py:163-173,191-197,224-229 drive it.

    python tests/golden/make_prednet_golden.py            (needs /root/reference; rewrites the fixtures)

Each fixture holds the recipe (stack, frame shape, seeds, bias kind) and, from the reference run:
  p0      Model.predict(x)[0, 0]                      (input independent, compress.py:197)
  next1   Model.predict([frame, 0])[:, 1]             (compress.py:224-229)
  next2   Model.predict([next1, 0])[:, 1]             (the fed-back prediction of the following step, :222)
These pin oracle/prednet_oracle.py (<= 1e-6, tests/test_prednet_golden.py) and the tcgen05 path (<= 6e-3,
tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.keras_shim import ReferencePredNet          # noqa: E402
from tezip_b200 import synth                             # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prednet")

CASES = {
    # name: (stack, Hp, Wp, B, frame seed, bias, weight seed)
    "tiny_uniform": ((3, 16, 32, 64), 24, 40, 3, 3, "uniform", 7),
    "tiny_zeros": ((3, 16, 32, 64), 24, 40, 3, 4, "zeros", 7),
    "mono": ((1, 16, 32, 64), 16, 24, 2, 5, "uniform", 7),
    "odd_stack": ((3, 8, 24, 40), 16, 32, 2, 6, "uniform", 11),
    "full_128x160": ((3, 48, 96, 192), 128, 160, 2, 1, "uniform", 7),       # train.py:51, the BASELINE frame shape
}


def main():
    os.makedirs(HERE, exist_ok=True)
    for name, (stack, Hp, Wp, B, seed, bias, wseed) in CASES.items():
        ws = synth.make_weights(stack, bias=bias, seed=wseed)
        net = ReferencePredNet(ws, stack, stack, Hp, Wp)
        fr = synth.make_frames(B, Hp, Wp, stack[0], seed=seed).astype(np.float32) / 255
        x = np.stack([fr, np.zeros_like(fr)], axis=1)
        out = net.predict(x, 10)
        p0, next1 = out[0, 0], out[:, 1]
        assert all(np.array_equal(out[b, 0], p0) for b in range(B))
        next2 = net.next(next1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), stack=np.array(stack), Hp=Hp, Wp=Wp, B=B, seed=seed,
                            bias=bias, wseed=wseed, p0=p0, next1=next1, next2=next2)
        print(name, "p0 [%.4f, %.4f]" % (p0.min(), p0.max()), "next1 mean %.4f" % next1.mean(),
              "next2 mean %.4f" % next2.mean())


if __name__ == "__main__":
    main()
