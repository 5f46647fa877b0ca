"""Loads the CPU emulation of the GPU zstd frame writer (tests/zstd_emu.cpp: the per-thread bodies of
tezip_b200/csrc/tz_zstd_core.h run in plain loops).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "zstd_emu.cpp")
CORE = os.path.join(HERE, "..", "tezip_b200", "csrc", "tz_zstd_core.h")
OUT = os.path.join(HERE, "_build", "libzstd_emu.so")
_emu = None


def load():
    global _emu
    if _emu is None:
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            tmp = OUT + ".tmp.%d" % os.getpid()
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", tmp, SRC])
            os.replace(tmp, OUT)
        e = ctypes.CDLL(OUT)
        vp, ull = ctypes.c_void_p, ctypes.c_ulonglong
        e.emu_bound.restype, e.emu_bound.argtypes = ull, [ull]
        e.emu_hist.restype, e.emu_hist.argtypes = None, [vp, ull, vp, vp]
        e.emu_encode.restype, e.emu_encode.argtypes = ull, [vp, ull, vp, vp, ctypes.c_uint, vp, vp]
        e.emu_decode.restype, e.emu_decode.argtypes = ctypes.c_int, [vp, vp, ull, vp, vp]
        _emu = e
    return _emu


def compress(a):
    """The frame tezip_b200.zstd_frames.compress_device writes for the bytes of `a`, computed on the CPU."""
    from tezip_b200 import zstd_frames as zf
    emu = load()
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    n = raw.size
    if n == 0:
        return zf.empty_frame()
    hist = np.zeros(256, np.uint32)
    uniform = np.zeros(-(-n // zf.BLOCK), np.int32)
    emu.emu_hist(raw.ctypes.data, n, hist.ctypes.data, uniform.ctypes.data)
    ct, tree = zf.huffman_tables(hist)
    tb = np.frombuffer(tree + b"\0", np.uint8).copy()
    out = np.zeros(emu.emu_bound(n) // 4, np.uint32)
    size = emu.emu_encode(raw.ctypes.data, n, ct.ctypes.data, tb.ctypes.data, len(tree), uniform.ctypes.data,
                          out.ctypes.data)
    return out.view(np.uint8)[:size].tobytes()


def decompress(frame):
    """What tezip_b200.zstd_frames.decompress_device returns for `frame`, computed on the CPU; None if the frame is
    outside the subset the kernels read."""
    from tezip_b200 import zstd_frames as zf
    parsed = zf.parse_frame(frame)
    if parsed is None:
        return None
    content, blocks, tables = parsed
    out = np.zeros(max(content, 1), np.uint8)
    src = np.zeros(len(frame) + 8, np.uint8)                  # (the stream decoder reads aligned 32-bit words)
    src[:len(frame)] = np.frombuffer(frame, np.uint8)
    blocks = np.ascontiguousarray(blocks)
    tables = np.ascontiguousarray(tables)
    err = load().emu_decode(src.ctypes.data, blocks.ctypes.data, len(blocks), tables.ctypes.data, out.ctypes.data)
    if err:
        raise RuntimeError("corrupt zstd frame (Huffman stream error %d)" % err)
    return out[:content]


def cases():
    """name -> array: the shapes of data the container holds, and the corners of the frame layout."""
    rng = np.random.default_rng(11)
    key = np.zeros((12, 40, 56, 3), np.uint8)
    key[::4] = rng.integers(0, 256, key[::4].shape)
    return {
        "empty": np.zeros(0, np.uint8),
        "one_byte": np.array([9], np.uint8),
        "zeros": np.zeros(300000, np.uint8),
        "key_plane": key,
        "ranks_int16": rng.geometric(0.15, 300001).clip(0, 700).astype(np.int16),
        "ranks_int32": rng.geometric(0.01, 70001).astype(np.int32),
        "uniform256": rng.integers(0, 256, 200000).astype(np.uint8),
        "two_values": (rng.integers(0, 2, 150000) * 200).astype(np.uint8),
        "below_min": rng.integers(0, 5, 1023).astype(np.uint8),
        "at_min": rng.integers(0, 5, 1024).astype(np.uint8),
        "exact_block": rng.integers(0, 9, 131072).astype(np.uint8),
        "block_plus_one": rng.integers(0, 9, 131073).astype(np.uint8),
        "mixed_blocks": np.concatenate([rng.integers(0, 9, 131072), np.zeros(131072, np.int64),
                                        rng.integers(0, 256, 131072), rng.integers(0, 3, 1500)]).astype(np.uint8),
        "deep_tree": np.repeat(np.arange(40, dtype=np.uint8), (1.5 ** np.arange(40)).astype(int) + 1)[:600000],
        "high_symbols": (255 - rng.geometric(0.2, 200000).clip(0, 255)).astype(np.uint8),
    }
