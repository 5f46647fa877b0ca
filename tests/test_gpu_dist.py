"""Sharding on one GPU: encoding a sequence as window-aligned shards with the halo / histogram exchange
(tezip_b200/dist.py protocol, emulated in-process) gives the same stream as encoding it in one piece, and every
shard decodes independently."""
import numpy as np
import pytest

from helpers import TINY, oracle_net, gpu_net
from tezip_b200 import synth

pytestmark = pytest.mark.gpu


class _FakeComm:
    """Runs the ranks one after the other: pass 1 records last_x / histograms, pass 2 replays the reductions."""
    def __init__(self, world):
        self.world, self.rank, self.phase = world, 0, 0
        self.last_x = [0] * world
        self.hists = {}
        self.calls = 0

    def exchange_last_x(self, x_last):
        if self.phase == 0:
            self.last_x[self.rank] = x_last.clone()          # device int32[1], as ShardComm keeps it
        return (self.rank > 0), (self.last_x[self.rank - 1] if self.rank > 0 else None)

    def reduce_hist(self, t):
        key = self.calls
        self.calls += 1
        if self.phase == 0:
            self.hists.setdefault(key, []).append(t.clone())
        else:
            t.copy_(sum(self.hists[key]))


@pytest.mark.parametrize("mode,bound", [("abs", [0.0]), ("abs", [2.0])])
def test_sharded_encode_equals_unsharded(cuda_lib, mode, bound):
    import torch
    from tezip_b200 import codec
    from tezip_b200.dist import shard_ranges
    stack, H, W, nt, Wn, world = TINY, 24, 40, 23, 4, 3
    _o, ws = oracle_net(stack)
    net = gpu_net(stack, ws, 24, 40, max_batch=8)
    frames = synth.make_frames(nt, H, W, 3, seed=31)
    dev = torch.device("cuda", 0)
    fr = torch.from_numpy(frames).to(dev)
    whole = codec.encode_frames(fr, net, 0, Wn, None, mode, bound, True)
    ranges = shard_ranges(nt, 0, Wn, world)
    comm = _FakeComm(world)
    encs = None
    for phase in (0, 1):
        comm.phase = phase
        encs = []
        for r, (a, b) in enumerate(ranges):
            comm.rank, comm.calls = r, 0
            encs.append(codec.encode_frames(fr[a:b].contiguous(), net, 0, Wn, None, mode, bound, True, comm=comm))
    body = torch.cat([e.body for e in encs]).cpu().numpy()
    assert all(np.array_equal(e.table, whole.table) for e in encs)
    assert np.array_equal(body, whole.body.cpu().numpy())
    assert np.array_equal(torch.cat([e.key_plane for e in encs]).cpu().numpy(), whole.key_plane.cpu().numpy())
    # every shard decodes on its own: x restarts at 0 at its first (key) frame
    outs = []
    for r, ((a, b), e) in enumerate(zip(ranges, encs)):
        out, _ = codec.decode_arrays(e.key_plane, e.body, e.table, e.shape, 0, net,
                                     first_mode=0 if r == 0 else 1, first_x=0)
        outs.append(out)
    dec = torch.cat(outs).cpu().numpy()
    if bound == [0.0]:
        assert np.array_equal(dec, frames)
    else:
        assert np.abs(dec.astype(int) - frames.astype(int)).max() <= 2


def test_one_frame_trailing_shard_and_zero_key_frame(cuda_lib):
    """Advisor cases: 11 frames, window 5, 3 ranks -> the last rank owns ONE frame (a lone key) and must encode it;
    an all-zero key frame is refused at compress time (the decoder could not find it, decompress.py:123-127)."""
    import torch
    from tezip_b200 import codec
    from tezip_b200._lib import TezipError
    from tezip_b200.dist import shard_ranges
    stack, H, W, nt, Wn, world = TINY, 24, 40, 11, 5, 3
    _o, ws = oracle_net(stack)
    net = gpu_net(stack, ws, 24, 40, max_batch=4)
    frames = synth.make_frames(nt, H, W, 3, seed=31)
    fr = torch.from_numpy(frames).cuda()
    whole = codec.encode_frames(fr, net, 0, Wn, None, "abs", [0.0], True)
    ranges = shard_ranges(nt, 0, Wn, world)
    assert ranges[-1] == (10, 11)
    comm = _FakeComm(world)
    for phase in (0, 1):
        comm.phase = phase
        encs = []
        for r, (a, b) in enumerate(ranges):
            comm.rank, comm.calls = r, 0
            encs.append(codec.encode_frames(fr[a:b].contiguous(), net, 0, Wn, None, "abs", [0.0], True, comm=comm))
    assert np.array_equal(torch.cat([e.body for e in encs]).cpu().numpy(), whole.body.cpu().numpy())
    black = fr.clone()
    black[5] = 0                                            # frame 5 is a key frame of the schedule
    with pytest.raises(TezipError):
        codec.encode_frames(black, net, 0, Wn, None, "abs", [0.0], True)
    black[5] = fr[5]
    black[6] = 0                                            # a black NON-key frame is fine
    enc = codec.encode_frames(black, net, 0, Wn, None, "abs", [0.0], True)
    out, _ = codec.decode_arrays(enc.key_plane, enc.body, enc.table, enc.shape, 0, net)
    assert torch.equal(out, black)
    net.close()
