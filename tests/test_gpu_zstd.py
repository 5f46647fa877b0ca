"""The GPU zstd frame writer on a B200: the kernels must write byte for byte what their CPU emulation writes (same
per-thread bodies, tests/zstd_emu.cpp), libzstd must decode the frames to the source, and a container written with
TEZIP_ZSTD_LEVEL=gpu must decompress like one written by libzstd level 9 (compress.py:276,398; decompress.py:89,98)."""
import os

import numpy as np
import pytest

import zstd_emu
from helpers import TINY

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(zstd_emu.cases()))
def test_device_frame_equals_emulation(cuda_lib, name):
    import torch
    from tezip_b200 import container, zstd_frames as zf
    a = zstd_emu.cases()[name]
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    l0 = cuda_lib.tz_launch_count()
    frame = zf.compress_device(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    assert raw.size == 0 or cuda_lib.tz_launch_count() > l0
    assert np.array_equal(container.zstd_decompress(frame), raw)
    assert frame == zstd_emu.compress(a)
    back = zf.decompress_device(frame, "cuda:0")             # the device decoder reads it back
    assert back is not None and back.dtype == torch.uint8 and np.array_equal(back.cpu().numpy(), raw)


def test_device_frame_large_stream(cuda_lib):
    """A stream of the bench's size class (64 MB of int16 ranks with a zero high byte): decodes to the source; rate and
    ratio are bench.py's business."""
    import torch
    from tezip_b200 import container, zstd_frames as zf
    g = torch.Generator(device="cuda").manual_seed(5)
    t = (torch.empty(32 << 20, device="cuda").exponential_(0.2, generator=g)).clamp_(0, 300).to(torch.int16)
    frame = zf.compress_device(t)
    assert len(frame) < t.numel()           # < 8 bits per int16 code
    assert np.array_equal(container.zstd_decompress(frame).view("<i2"), t.cpu().numpy())
    assert torch.equal(zf.decompress_device(frame, t.device).view(torch.int16), t)


def test_device_decoder_leaves_libzstd_frames_alone_and_reports_corruption(cuda_lib):
    import torch
    from tezip_b200 import container, zstd_frames as zf
    a = (np.arange(300000) % 251).astype(np.uint8)
    assert zf.decompress_device(container.zstd_compress(a), "cuda:0") is None
    src = torch.from_numpy(np.random.default_rng(4).geometric(0.2, 200000).clip(0, 255).astype(np.uint8)).cuda()
    frame = bytearray(zf.compress_device(src))
    _content, blocks, _tables = zf.parse_frame(bytes(frame))
    frame[int(blocks[0]["src_off"]) + int(blocks[0]["stream_bytes"][0]) - 1] = 0
    with pytest.raises(RuntimeError):
        zf.decompress_device(bytes(frame), "cuda:0")


def test_container_written_on_gpu_round_trips(cuda_lib, tmp_path, monkeypatch):
    import torch
    from tezip_b200 import codec, container, synth
    from tezip_b200.prednet import PredNet
    stack, H, W, nt = TINY, 24, 40, 13
    net = PredNet(stack, stack, weights=synth.make_weights(stack, bias="uniform", seed=3), input_hw=(H, W),
                  max_batch=4, device=0)
    frames = synth.make_frames(nt, H, W, 3, seed=4)
    enc = codec.encode_frames(torch.from_numpy(frames).cuda(), net, 0, 4, None, "abs", [0.0], True)
    payload = enc.payload()
    names = ["f%02d.png" % i for i in range(nt)]
    ref_dir, gpu_dir = str(tmp_path / "ref"), str(tmp_path / "gpu")
    container.write_container(ref_dir, names, True, enc.key_plane.cpu().numpy(), payload, workers=0)
    tail = codec.pack_payload(enc.body[:0].cpu().numpy(), enc.table, enc.shape, enc.p)
    container.write_container_device(gpu_dir, names, True, enc.key_plane, enc.body, tail)
    a, b = container.read_container(ref_dir), container.read_container(gpu_dir)
    assert a[0] == b[0] and a[1] == b[1]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    body, table, shape, p = codec.parse_payload(b[3])
    out, _ = codec.decode_arrays(torch.from_numpy(np.ascontiguousarray(b[2]).reshape(shape[1:])).cuda(),
                                 torch.from_numpy(np.ascontiguousarray(body)).cuda(), table, shape, p, net)
    assert np.array_equal(out.cpu().numpy(), frames)
    net.close()
