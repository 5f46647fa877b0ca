"""A/B of the host-side pipeline on one GPU (development aid): device-resident compress with the host reading every
step's table before the next step (defer=False) vs after the next step has been queued (defer=True), and the same for
the streaming host-buffer path.  BASELINE config 2."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tezip_b200 import synth, codec
from tezip_b200.prednet import PredNet
STACK = (3, 48, 96, 192)
nt, H, W, C = 1000, 128, 160, 3
dev = torch.device("cuda", 0)
ws = synth.make_weights(STACK, bias="uniform", seed=7)
net = PredNet(STACK, STACK, weights=ws, input_hw=(H, W), max_batch=100)
fh = torch.from_numpy(synth.make_frames(nt, H, W, C, seed=1)).pin_memory()
fd = fh.to(dev)
sets = [(torch.empty_like(fh).pin_memory(), torch.empty(fh.numel(), dtype=torch.int16).pin_memory()) for _ in range(2)]
prev = [None]
seq = [0]
keep = [None, None]


def settle(e):
    if prev[0] is not None:
        prev[0].finalize()
    prev[0] = e


def dev_sync():
    codec.encode_frames(fd, net, 0, 10, None, "abs", [2.0], True)


def dev_defer():
    settle(codec.encode_frames(fd, net, 0, 10, None, "abs", [2.0], True, defer=True))


def host(defer):
    kh, bh = sets[seq[0] & 1]
    seq[0] += 1
    e = codec.encode_frames_host(fh, net, 0, 10, None, "abs", [2.0], kh, bh, True, wait_copies=False, defer=defer)
    keep[seq[0] & 1] = e
    if defer:
        settle(e)


def timed(fn, steps=8, warm=2):
    for _ in range(warm):
        fn()
    settle(None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    settle(None)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for rnd in range(2):
    print("round %d: dev sync %.3f ms | dev defer %.3f ms | host stream %.3f ms | host stream defer %.3f ms" % (
        rnd, timed(dev_sync), timed(dev_defer), timed(lambda: host(False)), timed(lambda: host(True))), flush=True)
