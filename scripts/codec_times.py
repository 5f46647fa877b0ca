"""Times the codec kernels of the BASELINE config-2 shape on one GPU (development aid; bench.py reports the same)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tezip_b200 import synth, ops, codec, _lib
from tezip_b200.prednet import PredNet
STACK = (3, 48, 96, 192)
nt, H, W, C = 1000, 128, 160, 3
dev = torch.device("cuda", 0)
ws = synth.make_weights(STACK, bias="uniform", seed=7)
net = PredNet(STACK, STACK, weights=ws, input_hw=(H, W), max_batch=100)
fr = torch.from_numpy(synth.make_frames(nt, H, W, C, seed=1)).to(dev)
enc = codec.encode_frames(fr, net, 0, 10, None, "abs", [2.0], True, keep_pool=True)
pool, slot = enc.pool, torch.from_numpy(enc.pred_slot).to(dev)
apply_t = torch.from_numpy((enc.pred_slot >= 0).astype(np.uint8)).to(dev)
x = torch.empty((nt, H, W, C), dtype=torch.int16, device=dev)
h = torch.zeros(_lib.TZ_HIST_BINS + 2, dtype=torch.int64, device=dev)


def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def fused():
    h[-1:].zero_()
    ops.encode_lossy(fr, pool, slot, apply_t, "abs", [2.0], h[:4096], h[-2:-1], h[-1:].view(torch.int32), x=x)


print(os.environ.get("TAG", ""), "fused_lossy %.4f ms" % ev(fused), "| residual+eb %.4f" % ev(lambda: (ops.residual(fr, pool, slot, out=x), ops.error_bound(fr, x, apply_t, "abs", [2.0]))))

# the stream kernels (delta + histogram, delta + rank map, fused lossless passes, decoder)
N = nt * H * W * C
ops.residual(fr, pool, slot, out=x)
ops.error_bound(fr, x, apply_t, "abs", [2.0])
hist = torch.zeros(_lib.TZ_HIST_BINS, dtype=torch.int64, device=dev)
ovf = torch.zeros(1, dtype=torch.int64, device=dev)
lut = torch.from_numpy(ops.encode_lut(enc.table)).to(dev)
obuf = torch.empty(N, dtype=torch.int16, device=dev)
lut_d = torch.from_numpy(ops.decode_lut(enc.table)).to(dev)
t = {
    "delta_hist": ev(lambda: ops.finding_difference_hist(x, hist, ovf)),
    "delta_rank": ev(lambda: ops.finding_difference_rank(x, lut, out=obuf)),
    "lossless_hist": ev(lambda: ops.encode_lossless(fr, pool, slot, 0, hist=hist, overflow=ovf)),
    "lossless_rank": ev(lambda: ops.encode_lossless(fr, pool, slot, 1, lut=lut, out=obuf)),
    "reconstruct": ev(lambda: ops.reconstruct(enc.body, (nt, H, W, C), H, W, len(enc.table), lut_d, pool, slot,
                                              enc.key_plane)),
}
bps = {"delta_hist": 2, "delta_rank": 4, "lossless_hist": 5, "lossless_rank": 7, "reconstruct": 7}
print(os.environ.get("TAG", ""), " ".join("%s %.4f ms (%.0f GB/s)" % (k, v, bps[k] * N / v / 1e6) for k, v in t.items()))
