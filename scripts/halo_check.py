"""Experiment: tensor-core path with / without the halo mode vs the fp32 oracle (prints max abs errors)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tezip_b200 import synth
from tezip_b200.prednet import PredNet
from oracle.prednet_oracle import PredNetOracle
for stack, H, W in (((3, 16, 32, 64), 32, 64), ((3, 48, 96, 192), 32, 48), ((3, 48, 96, 192), 128, 160)):
    ws = synth.make_weights(stack, bias="uniform", seed=7)
    onet = PredNetOracle(ws, stack, stack)
    fr = synth.make_frames(3, H, W, 3, seed=4).astype(np.float32) / 255
    ref = onet.next(fr)
    net = PredNet(stack, stack, weights=ws, input_hw=(H, W), max_batch=4)
    got = net.next(torch.from_numpy(fr).cuda()).cpu().numpy()
    print(os.environ.get("TZ_HALO", "1"), stack, H, W, "max err %.3e" % np.abs(got - ref).max(), flush=True)
    net.close()
