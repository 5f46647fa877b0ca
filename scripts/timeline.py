"""Kernel timeline of ONE compress step (codec.encode_frames, BASELINE config 2) from the CUPTI activity records that
torch.profiler collects: where the GPU idles between kernels.  Diagnostic, not a bench (profiling adds overhead)."""
import os
import sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tezip_b200 import synth, codec          # noqa: E402
from tezip_b200.prednet import PredNet       # noqa: E402

STACK = (3, 48, 96, 192)
nt, W = 1000, 10
ws = synth.make_weights(STACK, bias="uniform", seed=7)
frames = torch.from_numpy(synth.make_frames(nt, 128, 160, 3, seed=1)).cuda()
net = PredNet(STACK, STACK, weights=ws, input_hw=(128, 160), max_batch=100)


def step():
    return codec.encode_frames(frames, net, 0, W, None, "abs", [2.0], True)


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
print("%d device activities, span %.3f ms, sum of durations %.3f ms" % (len(ev), (t1 - t0) / 1e3, busy / 1e3))
gaps = []
end = ev[0].time_range.end
for a, b in zip(ev, ev[1:]):
    g = b.time_range.start - end
    if g > 0:
        gaps.append((g, a.name[:50], b.name[:50], (b.time_range.start - t0) / 1e3))
    end = max(end, b.time_range.end)
print("idle total %.3f ms in %d gaps; gaps > 3 us: %.3f ms" % (sum(g for g, *_ in gaps) / 1e3, len(gaps),
                                                               sum(g for g, *_ in gaps if g > 3) / 1e3))
for g, a, b, at in sorted(gaps, reverse=True)[:25]:
    print("%8.1f us at %7.3f ms  after %-50s before %s" % (g, at, a, b))
by = {}
for e in ev:
    k = e.name[:60]
    by[k] = by.get(k, 0) + (e.time_range.end - e.time_range.start)
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:14]:
    print("%9.1f us  %s" % (v, k))
