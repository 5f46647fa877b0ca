"""Where one compress step's time goes besides the PredNet kernels: CUDA events at the phase boundaries of
codec.encode_frames (static windows) + host clocks of the same phases.  Diagnostic, not a bench."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tezip_b200 import synth, codec          # noqa: E402
from tezip_b200.prednet import PredNet       # noqa: E402

STACK = (3, 48, 96, 192)
nt, W = 1000, 10
ws = synth.make_weights(STACK, bias="uniform", seed=7)
frames = torch.from_numpy(synth.make_frames(nt, 128, 160, 3, seed=1)).cuda()
net = PredNet(STACK, STACK, weights=ws, input_hw=(128, 160), max_batch=100)
dev = frames.device
st = torch.cuda.current_stream(dev)


def ev():
    return st.record_event(torch.cuda.Event(enable_timing=True))


def one():
    t = [time.perf_counter()]
    e = [ev()]
    plan = codec.cached_plan(nt, 0, codec.swp_keys(nt, 0, W))
    pool = torch.empty((plan.n_slots, 128, 160, 3), dtype=torch.float32, device=dev)
    t.append(time.perf_counter()); e.append(ev())
    staged = codec.stage_plan(frames, plan.keys, plan.pred_slot, plan.apply_eb, None, plan)
    t.append(time.perf_counter()); e.append(ev())
    codec.run_plan(net, frames, plan, pool)
    t.append(time.perf_counter()); e.append(ev())
    enc = codec.encode_with_pool(frames, pool, plan.pred_slot, plan.apply_eb, plan.keys, 0, "abs", [2.0], True, staged=staged)
    t.append(time.perf_counter()); e.append(ev())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return np.diff(t) * 1e3, [e[i].elapsed_time(e[i + 1]) for i in range(len(e) - 1)]


for _ in range(3):
    one()
H, G = [], []
for _ in range(5):
    h, g = one()
    H.append(h); G.append(g)
H, G = np.mean(H, 0), np.mean(G, 0)
names = ["plan", "stage_plan", "run_plan", "encode_with_pool", "final sync"]
for i, n in enumerate(names):
    print("%-18s host %.3f ms   gpu(events) %s" % (n, H[i], "%.3f ms" % G[i] if i < len(G) else "-"))
print("total host %.3f  gpu %.3f" % (H.sum(), sum(G)))
