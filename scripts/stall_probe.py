"""Where does the one-off host stall of the streaming host-buffer calls come from?  Per call: host time, cudaMalloc
count and reserved bytes of the caching allocator."""
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
sets = [(torch.empty_like(fh).pin_memory(), torch.empty(fh.numel(), dtype=torch.int16).pin_memory()) for _ in range(2)]
prev, keep, seq = [None], [None, None], [0]


def stats():
    s = torch.cuda.memory_stats(dev)
    return "mallocs %d retries %d reserved %.0f MB" % (s.get("num_device_alloc", -1), s.get("num_alloc_retries", -1),
                                                       s["reserved_bytes.all.current"] / 1e6)


def host():
    kh, bh = sets[seq[0] & 1]
    seq[0] += 1
    e = codec.encode_frames_host(fh, net, 0, 10, None, "abs", [2.0], kh, bh, True, wait_copies=False, defer=True)
    keep[seq[0] & 1] = e
    if prev[0] is not None:
        prev[0].finalize()
    prev[0] = e


for rnd in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(8):
        t1 = time.perf_counter()
        host()
        print("round %d call %d: host %.2f ms (at %.2f) %s" % (rnd, i, (time.perf_counter() - t1) * 1e3,
                                                              (time.perf_counter() - t0) * 1e3, stats()), flush=True)
    prev[0].finalize()
    prev[0] = None
    torch.cuda.synchronize()
    print("round %d done at %.2f ms" % (rnd, (time.perf_counter() - t0) * 1e3), flush=True)
