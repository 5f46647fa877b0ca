"""PredNet on the GPU vs the fp32 oracle (oracle/prednet_oracle.py).

Tolerances (stated per north_star):
  * fp32 direct kernels: |pred - oracle| <= 2e-5 (same arithmetic, different summation order)
  * tcgen05 path (fp16 operands, fp32 accumulation): |pred - oracle| <= 6e-3 on predictions in [0, 1]
Bitwise properties: predictions do not depend on the batch size or on which frames share a batch."""
import numpy as np
import pytest

from helpers import TINY, FULL, oracle_net, gpu_net
from tezip_b200 import synth

pytestmark = pytest.mark.gpu

TOL_DIRECT = 2e-5
TOL_TC = 6e-3


def _inputs(n, H, W, Hp, Wp, seed=4, C=3):
    fr = synth.make_frames(n, H, W, C, seed=seed).astype(np.float32) / 255
    x = np.zeros((n, Hp, Wp, C), np.float32)
    x[:, :H, :W] = fr
    return x


@pytest.mark.parametrize("stack,H,W", [(TINY, 20, 28), (TINY, 32, 64), (FULL, 32, 48)])
@pytest.mark.parametrize("bias", ["uniform", "zeros"])
def test_direct_matches_oracle(cuda_lib, stack, H, W, bias):
    import torch
    Hp, Wp = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    onet, ws = oracle_net(stack, bias=bias)
    net = gpu_net(stack, ws, Hp, Wp, max_batch=4, fp32_direct=True)
    x = _inputs(3, H, W, Hp, Wp)
    ref = onet.next(x)
    got = net.next(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(got - ref).max() <= TOL_DIRECT
    assert np.abs(net.p0().cpu().numpy() - onet.p0(Hp, Wp)).max() <= TOL_DIRECT
    # chained predictions (prediction fed back, padded border not re-zeroed: compress.py:222)
    ref2 = onet.next(ref)
    got2 = net.next(torch.from_numpy(got).cuda()).cpu().numpy()
    assert np.abs(got2 - ref2).max() <= 5 * TOL_DIRECT
    net.close()


@pytest.mark.parametrize("stack,H,W", [(TINY, 20, 28), (TINY, 32, 64), (TINY, 128, 128), (FULL, 32, 48),
                                       (FULL, 128, 160)])
def test_tc_matches_oracle(cuda_lib, stack, H, W):
    import torch
    Hp, Wp = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    onet, ws = oracle_net(stack)
    net = gpu_net(stack, ws, Hp, Wp, max_batch=4)
    x = _inputs(3, H, W, Hp, Wp)
    ref = onet.next(x)
    got = net.next(torch.from_numpy(x).cuda()).cpu().numpy()
    err = np.abs(got - ref).max()
    assert err <= TOL_TC, err
    assert np.abs(net.p0().cpu().numpy() - onet.p0(Hp, Wp)).max() <= TOL_DIRECT
    net.close()


@pytest.mark.parametrize("direct", [True, False])
def test_batch_invariance_bitwise(cuda_lib, direct):
    """compress and decompress may batch differently: the same frame must give the same bits."""
    import torch
    stack, H, W = TINY, 24, 40
    _onet, ws = oracle_net(stack)
    net = gpu_net(stack, ws, H, W, max_batch=7, fp32_direct=direct)
    x = torch.from_numpy(_inputs(7, H, W, H, W, seed=9)).cuda()
    full = net.next(x).cpu().numpy()
    for idx in ([0], [6], [3, 1], [5, 4, 2, 0, 6]):
        sub = net.next(x[idx].contiguous()).cpu().numpy()
        assert np.array_equal(sub, full[idx])
    net.close()


def test_predict_protocol(cuda_lib):
    """keras-style predict([frame, zeros]) mirrors the reference call sites."""
    stack, H, W = TINY, 16, 24
    onet, ws = oracle_net(stack)
    net = gpu_net(stack, ws, H, W, max_batch=2, fp32_direct=True)
    x = np.zeros((1, 2, H, W, 3))
    x[0, 0] = _inputs(1, H, W, H, W)[0]
    got, ref = net.predict(x, 10), onet.predict(x, 10)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= TOL_DIRECT
    got1 = net.predict(x[:, :1], 10)
    assert np.abs(got1 - onet.predict(x[:, :1], 10)).max() <= TOL_DIRECT
    net.close()


@pytest.mark.parametrize("direct", [True, False])
def test_chained_steps_bitwise(cuda_lib, direct):
    """next_chained (input = the previous prediction, compress.py:222-229) equals next() on that tensor bit for
    bit, with a shrinking batch (windows that end), and refuses what it cannot honour."""
    import torch
    from tezip_b200._lib import TezipError
    stack, H, W = TINY, 24, 40
    _onet, ws = oracle_net(stack)
    net = gpu_net(stack, ws, H, W, max_batch=6, fp32_direct=direct)
    x = torch.from_numpy(_inputs(6, H, W, H, W, seed=13)).cuda()
    with pytest.raises(TezipError):
        net.next_chained(torch.empty_like(x))                   # nothing to chain from yet
    ref1 = net.next(x)
    ref2 = net.next(ref1)                                       # plain: 6 frames
    ref3 = net.next(ref2[:4].contiguous())                      # plain: the first 4 only
    a1 = net.next(x)
    a2 = net.next_chained(torch.empty_like(x))
    a3 = net.next_chained(torch.empty_like(x[:4]))
    assert torch.equal(a1, ref1) and torch.equal(a2, ref2) and torch.equal(a3, ref3)
    with pytest.raises(TezipError):
        net.next_chained(torch.empty_like(x))                   # 6 > 4: the batch may not grow
    with pytest.raises(TezipError):
        net.next_chained(a3)                                    # out is the previous prediction
    net.close()


@pytest.mark.parametrize("stack,H,W,B", [(FULL, 256, 320, 2),            # many tiles per image in both directions
                                         (FULL, 512, 512, 2),            # 16x the bench frame
                                         ((1, 16, 32, 64), 64, 96, 5),   # one channel (SURVEY 8(d) config 4 family)
                                         ((3, 24, 40), 48, 72, 7),       # three layers, widths not multiples of 16
                                         ((3, 48, 96, 192), 128, 160, 37),    # odd batch: ragged last CTA pair
                                         ((3, 32), 40, 56, 3),            # two layers: layer 1 is the top of the split
                                         ((1, 48, 96, 192), 128, 128, 2),     # the config-4 net on a small frame
                                         ((2, 16, 32), 32, 48, 3)])       # S_0 = 2: generic layer-0 kernels (no split)
def test_tc_matches_oracle_more_shapes(cuda_lib, stack, H, W, B):
    import torch
    onet, ws = oracle_net(stack)
    net = gpu_net(stack, ws, H, W, max_batch=B)
    x = _inputs(B, H, W, H, W, seed=21, C=stack[0])
    pred1 = net.next(torch.from_numpy(x).cuda())            # kept alive: the chained step's input
    got = pred1.cpu().numpy()
    check = sorted(set([0, B // 2, B - 1]))                 # the oracle is slow: spot-check frames across the batch
    ref = onet.next(x[check])
    err = np.abs(got[check] - ref).max()
    assert err <= TOL_TC, err
    # chained second step through next_chained
    got2 = net.next_chained(torch.empty_like(pred1)).cpu().numpy()
    ref2 = onet.next(ref)
    assert np.abs(got2[check] - ref2).max() <= TOL_TC
    net.close()
