"""Experiment: the 100 windows of one compress step as ONE group of 100 on one stream vs TWO groups of 50 on two
streams with two PredNet handles (persistent kernels of one stream can start on the SMs the other stream's kernel has
already left: tile-quantisation tails and kernel boundaries overlap).  Prints ms per 9-step chain, several rounds."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tezip_b200 import synth, ops
from tezip_b200.prednet import PredNet
STACK = (3, 48, 96, 192)
H, W = 128, 160
ws = synth.make_weights(STACK, bias="uniform", seed=7)
dev = torch.device("cuda", 0)
fr = torch.from_numpy(synth.make_frames(100, H, W, 3, seed=1)).to(dev)
x = ops.pad_normalize(fr, None, H, W)
net100 = PredNet(STACK, STACK, weights=ws, input_hw=(H, W), max_batch=100)
nets = [PredNet(STACK, STACK, weights=ws, input_hw=(H, W), max_batch=50) for _ in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
out1 = [torch.empty_like(x) for _ in range(2)]
out2 = [[torch.empty_like(x[:50]) for _ in range(2)] for _ in range(2)]


def chain(net, xin, bufs):
    net.next(xin, out=bufs[0])
    a, b = bufs
    for _ in range(8):
        net.next_chained(b)
        a, b = b, a


def one():
    chain(net100, x, out1)


def two():
    main = torch.cuda.current_stream(dev)
    ev = main.record_event()
    for g in range(2):
        streams[g].wait_event(ev)
        with torch.cuda.stream(streams[g]):
            chain(nets[g], x[50 * g:50 * g + 50], out2[g])
    for g in range(2):
        main.wait_stream(streams[g])


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for rnd in range(3):
    print("round %d: one group of 100: %.3f ms | two groups of 50 on two streams: %.3f ms" % (rnd, timed(one), timed(two)),
          flush=True)
# same predictions? (chained buffers: after 9 steps the last prediction sits in bufs[0])
torch.cuda.synchronize()
one(); two(); torch.cuda.synchronize()
print("bitwise equal:", bool(torch.equal(out1[0][:50], out2[0][0]) and torch.equal(out1[0][50:], out2[1][0])))
