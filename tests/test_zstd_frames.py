"""The GPU zstd frame writer, checked on the CPU: the host-side Huffman / FSE logic of tezip_b200/zstd_frames.py and
the per-thread kernel bodies (run by tests/zstd_emu.cpp) must give frames that libzstd -- the decoder behind the
reference's zstd.decompress, decompress.py:89,98 -- decodes to the source bytes.  tests/test_gpu_zstd.py then checks
that the kernels write the same bytes."""
import numpy as np
import pytest

import zstd_emu
from tezip_b200 import container, zstd_frames as zf


@pytest.mark.parametrize("name", sorted(zstd_emu.cases()))
def test_emulated_frame_decodes_with_libzstd(name):
    a = zstd_emu.cases()[name]
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    frame = zstd_emu.compress(a)
    assert frame[:4] == b"\x28\xb5\x2f\xfd"
    assert np.array_equal(container.zstd_decompress(frame), raw)
    if name in ("zeros", "key_plane", "ranks_int16", "two_values", "deep_tree"):
        assert len(frame) < raw.size // 2


def test_frame_layout_by_block_type():
    """RLE blocks cost 4 bytes, incompressible blocks 3 + their size, and a zero frame sequence almost nothing."""
    assert len(zstd_emu.compress(np.zeros(5 * 131072 + 7, np.uint8))) == 14 + 6 * 4
    rnd = np.random.default_rng(0).integers(0, 256, 2 * 131072).astype(np.uint8)
    assert len(zstd_emu.compress(rnd)) == 14 + 2 * (3 + 131072)
    assert zstd_emu.compress(np.zeros(0, np.uint8)) == zf.empty_frame()


def test_fuzz_against_libzstd():
    for it in range(120):
        r = np.random.default_rng(1000 + it)
        k, n = int(r.integers(2, 257)), int(r.integers(1024, 300000))
        kind = it % 4
        if kind == 0:
            p = r.random(k) ** r.integers(1, 12)
        elif kind == 1:
            p = 0.5 ** np.arange(k) + 1e-9 * r.random(k)
        elif kind == 2:
            p = np.ones(k)
        else:
            p = r.random(k)
            p[r.integers(0, k)] += 50
        vals = r.permutation(256)[:k]
        a = vals[r.choice(k, size=n, p=p / p.sum())].astype(np.uint8)
        if it % 7 == 0:
            a[:131072 * (it % 3)] = 7
        assert np.array_equal(container.zstd_decompress(zstd_emu.compress(a)), a), (it, k, n, kind)


def test_code_lengths_are_limited_and_complete():
    hist = np.zeros(256, np.int64)
    hist[:60] = (1.7 ** np.arange(60)).astype(np.int64) + 1          # a plain Huffman code would be 59 bits deep
    lens = zf.code_lengths(hist)
    assert lens.max() <= zf.MAX_BITS and (lens[:60] > 0).all() and (lens[60:] == 0).all()
    assert sum(2.0 ** -int(v) for v in lens if v) == 1.0
    w, ct = zf.weights_and_codes(lens)
    codes = sorted((int(ct[s]) >> 16, int(ct[s]) & 0xFFFF) for s in range(60))
    assert len(set(codes)) == 60
    assert zf.code_lengths(np.eye(1, 256, 5)[0]) is None              # one byte value only: RLE, not Huffman


def test_tree_description_forms():
    """4-bit direct weights below 129 symbols when that is shorter, FSE-compressed weights beyond; nothing when the
    weights cannot be described (256 equiprobable byte values: such blocks stay raw)."""
    hist = np.zeros(256, np.int64)
    hist[:4] = [8, 4, 2, 2]
    w, _ = zf.weights_and_codes(zf.code_lengths(hist))
    assert zf.tree_description(w)[0] == 127 + 3
    hist = (1000 * 0.97 ** np.arange(256)).astype(np.int64) + 1
    w, _ = zf.weights_and_codes(zf.code_lengths(hist))
    t = zf.tree_description(w)
    assert t[0] < 128 and len(t) == 1 + t[0]
    w, _ = zf.weights_and_codes(zf.code_lengths(np.ones(256, np.int64)))
    assert zf.tree_description(w) is None
    ct, tree = zf.huffman_tables(np.ones(256, np.int64))
    assert tree == b"" and not ct.any()


def test_gpu_writer_switch(monkeypatch):
    monkeypatch.delenv("TEZIP_ZSTD_LEVEL", raising=False)
    assert not container.gpu_writer() and container.container_level() == 9
    monkeypatch.setenv("TEZIP_ZSTD_LEVEL", "gpu")
    assert container.gpu_writer() and container.container_level() == 9
    monkeypatch.setenv("TEZIP_ZSTD_LEVEL", "3")
    assert not container.gpu_writer() and container.container_level() == 3


@pytest.mark.parametrize("name", sorted(zstd_emu.cases()))
def test_emulated_decoder_reads_the_frames(name):
    """Header walk (zstd_frames.parse_frame, tree descriptions included) + the per-stream decoder bodies."""
    a = zstd_emu.cases()[name]
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    got = zstd_emu.decompress(zstd_emu.compress(a))
    assert got is not None and np.array_equal(got, raw)


def test_weights_survive_fse_round_trip():
    for seed in range(200):
        r = np.random.default_rng(seed)
        k = int(r.integers(3, 257))
        hist = np.zeros(256, np.int64)
        hist[r.permutation(256)[:k]] = (r.random(k) ** int(r.integers(1, 10)) * 1e6).astype(np.int64) + 1
        w, ct = zf.weights_and_codes(zf.code_lengths(hist))
        last = int(np.nonzero(w)[0].max())
        packed = zf.fse_compress_weights(w[:last])
        if packed is not None:
            assert zf.fse_decompress_weights(packed) == [int(v) for v in w[:last]], seed
        tree = zf.tree_description(w)
        if tree is not None:
            table, used = zf.decode_table(tree)
            assert used == len(tree)
            for s in np.nonzero(w)[0]:                      # every code of the encoder's table decodes to its byte
                code, nb = int(ct[s]) & 0xFFFF, int(ct[s]) >> 16
                e = int(table[code << (zf.DLOG - nb)])
                assert e & 0xFF == s and e >> 8 == nb


def test_frames_outside_the_subset_are_left_to_libzstd():
    a = (np.arange(300000) % 251).astype(np.uint8)
    assert zf.parse_frame(container.zstd_compress(a)) is None                 # matches and sequences
    assert zf.parse_frame(container.zstd_compress(a, workers=2)) is None
    assert zf.parse_frame(b"") is None and zf.parse_frame(b"\x28\xb5\x2f\xfd") is None
    good = zstd_emu.compress(np.random.default_rng(3).geometric(0.2, 200000).clip(0, 255).astype(np.uint8))
    assert zf.parse_frame(good) is not None
    assert zf.parse_frame(good[:-1]) is None and zf.parse_frame(good + b"\0") is None   # truncated / trailing bytes


def test_corrupt_stream_is_reported():
    a = np.random.default_rng(4).geometric(0.2, 200000).clip(0, 255).astype(np.uint8)
    frame = bytearray(zstd_emu.compress(a))
    content, blocks, tables = zf.parse_frame(bytes(frame))
    end = int(blocks[0]["src_off"]) + int(blocks[0]["stream_bytes"][0])
    frame[end - 1] = 0                                                        # the first stream loses its end mark
    with pytest.raises(RuntimeError):
        zstd_emu.decompress(bytes(frame))
    start = int(blocks[0]["src_off"])
    frame[start:end - 1] = bytes(end - 1 - start)                             # all-zero payload: every symbol takes the
    frame[end - 1] = 1                                                        # longest code, the stream runs out of bits
    with pytest.raises(RuntimeError):
        zstd_emu.decompress(bytes(frame))


def test_frames_decode_with_a_second_zstd_build():
    """pyarrow bundles its own libzstd: a second decoder build reads the frames too (one-shot, size from the caller)."""
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("zstd"):
        pytest.skip("pyarrow without zstd")
    codec = pa.Codec("zstd")
    for name, a in zstd_emu.cases().items():
        raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        if raw.size == 0:
            continue
        out = codec.decompress(zstd_emu.compress(a), decompressed_size=raw.size)
        assert np.array_equal(np.frombuffer(out, np.uint8), raw), name
