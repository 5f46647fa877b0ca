"""Short profiling target: a few PredNet next() steps (B = 100 windows, 128x160x3, (3,48,96,192)) plus one codec
pass.  Used under ncu (B200_PROFILING.md); numbers printed here are never bench values."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tezip_b200 import synth, ops, codec          # noqa: E402
from tezip_b200.prednet import PredNet            # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
STACK = (3, 48, 96, 192)
ws = synth.make_weights(STACK, bias="uniform", seed=7)
frames = torch.from_numpy(synth.make_frames(B * 3, 128, 160, 3, seed=1)).cuda()
net = PredNet(STACK, STACK, weights=ws, input_hw=(128, 160), max_batch=B)
x = ops.pad_normalize(frames, torch.arange(B, dtype=torch.int32, device="cuda") * 3, 128, 160)
y = torch.empty_like(x)
for _ in range(steps):
    net.next(x, out=y)
    x, y = y, x
torch.cuda.synchronize()
enc = codec.encode_frames(frames, net, 0, 3, None, "abs", [2.0], True)
torch.cuda.synchronize()
print("ok", float(x.mean()))
