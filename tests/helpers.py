"""Shared builders for the parity tests (oracle side and GPU side see identical seeded inputs)."""
import numpy as np

from tezip_b200 import synth

TINY = (3, 16, 32, 64)          # small PredNet that still exercises every kernel shape family
FULL = (3, 48, 96, 192)         # train.py:51 architecture


def oracle_net(stack, bias="uniform", seed=7):
    from oracle.prednet_oracle import PredNetOracle
    ws = synth.make_weights(stack, bias=bias, seed=seed)
    return PredNetOracle(ws, stack, stack), ws


def gpu_net(stack, ws, Hp, Wp, max_batch=16, fp32_direct=False):
    from tezip_b200.prednet import PredNet
    return PredNet(stack, stack, weights=ws, input_hw=(Hp, Wp), max_batch=max_batch, fp32_direct=fp32_direct)


def pool_from_oracle(r, device, p=None):
    """Oracle per-frame predictions -> (pool tensor, pred_slot, apply_eb) for the GPU codec kernels:
    slot f+1 holds the oracle's prediction for frame f; window starts use -1."""
    import torch
    nt = r["preds"].shape[0]
    pool = torch.from_numpy(np.concatenate([r["preds"][:1], r["preds"]], axis=0)).to(device)
    pred_slot = np.arange(1, nt + 1, dtype=np.int32)
    apply_eb = np.ones(nt, np.uint8)
    if p is None:
        p = int(r["payload"][-1])          # the int16 trailer ends with p (compress.py:392); v2 callers pass p
    for wi, (first, n) in enumerate(r["windows"]):
        pred_slot[first] = -1
        apply_eb[first] = 0
        if p != 0 and wi == 0:
            apply_eb[first:first + n] = 0          # compress.py:315: no error_bound on the warm-up window
    return pool, pred_slot, apply_eb
