"""One pass of the GPU zstd frame writer and reader over a sequence-sized input (123 MB int16 stream whose high byte
is zero + 61 MB key plane with every tenth frame non-zero), timed with CUDA events; the command `ncu -k regex:zs_`
is pointed at (profiles/r2l_ncu_zstd.md).  Usage: python scripts/zstd_gpu_probe.py [rounds]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tezip_b200 import build, zstd_frames as zf  # noqa: E402

build.build()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n = 1000 * 128 * 160 * 3
stream = torch.empty(n, device=dev).exponential_(0.12, generator=g).clamp_(0, 250).to(torch.int16)
key = torch.zeros((1000, 128 * 160 * 3), dtype=torch.uint8, device=dev)
key[::10] = torch.randint(0, 256, (100, 128 * 160 * 3), dtype=torch.uint8, device=dev, generator=g)
for name, t in (("stream", stream), ("key_plane", key)):
    for r in range(rounds):
        marks = []
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out, size = zf.frame_device(t, marks)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        ms = {nm: a.elapsed_time(b) for nm, a, b in marks}
    frame = out[:size].cpu().numpy()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    dm = []
    back = zf.decompress_device(frame, dev, marks=dm)
    torch.cuda.synchronize(dev)
    dwall = time.perf_counter() - t0
    ms["decode"] = dm[0][1].elapsed_time(dm[0][2])
    nbytes = t.numel() * t.element_size()
    print("%s: %d -> %d bytes (%.3f), kernels %s ms, write wall %.2f ms; read back wall %.2f ms, identical %s" %
          (name, nbytes, size, size / nbytes, {k: round(v, 3) for k, v in ms.items()}, wall * 1e3, dwall * 1e3,
           bool(torch.equal(back, t.reshape(-1).view(torch.uint8)))))
