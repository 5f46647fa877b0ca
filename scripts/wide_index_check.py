"""Fused (frames + predictions) vs materialised-x delta stream of the 16-bit codec across the 2^31 / 2^32 element marks."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tezip_b200 import ops   # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2100
H = W = 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(3)
frames = torch.randint(0, 65536, (nt, H, W, 1), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
pool = torch.rand((8, H, W, 1), generator=g, device=dev, dtype=torch.float32)
slot = (torch.arange(nt, device=dev, dtype=torch.int32) % 9) - 1          # -1 (window start), 0..7
x = ops.residual(frames, pool, slot)
y_x = torch.empty(nt * H * W, dtype=torch.int32, device=dev)
y_f = torch.empty_like(y_x)
ops.encode16(frames, pool, slot, x, 1, lut=None, out=y_x)
ops.encode16(frames, pool, slot, None, 1, lut=None, out=y_f)
torch.cuda.synchronize()
bad = torch.nonzero(y_x != y_f).view(-1)
print("nt", nt, "elements", nt * H * W, "mismatches", bad.numel(), "first", bad[:6].tolist(),
      [hex(int(v)) for v in bad[:6].tolist()])
if bad.numel():
    i = int(bad[0])
    print("y_x", y_x[i - 1:i + 2].tolist(), "y_f", y_f[i - 1:i + 2].tolist(), "x", x.view(-1)[i - 2:i + 2].tolist())
if bad.numel():
    i = int(bad[0])
    wrong_prev = int(y_f[i]) + int(x.view(-1)[i])
    fr = frames.view(-1).to(torch.int32)
    pl = pool.view(-1)

    def q(v):
        return int((v * 65535.0).to(torch.float32).trunc().item())
    print("wrong prev", wrong_prev, "true", int(x.view(-1)[i - 1]))
    FE = H * W
    f = (i - 1) // FE
    for name, s, j in (("true", int(slot[f]), i - 1), ("slot f+1", int(slot[f + 1]), i - 1), ("slot f-1", int(slot[f - 1]), i - 1),
                       ("frames i", int(slot[f]), i), ("frames wrap", int(slot[f]), (i - 1) - 2 ** 31)):
        for poff in (FE - 1, -1, 0):
            if s >= 0:
                print(name, "poff", poff, q(pl[s * FE + poff]) - int(fr[j]))
