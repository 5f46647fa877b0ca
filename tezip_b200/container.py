"""The on-disk container: filename.txt, key_frame.dat, entropy.dat -- unchanged from
the reference (compress.py:133-136,271-278,394-400; decompress.py:48-56,87-103).

zstd: the reference calls python-zstd 1.4.5.1 `zstd.compress(bytes, 9)` / `zstd.decompress(bytes)`
(docs/index.rst:267), i.e. ONE frame that carries its content size.  Here libzstd.so.1 is driven through ctypes;
`workers > 0` uses the multithreaded single-frame encoder (ZSTD_c_nbWorkers), which the reference's decoder reads
as-is.
"""
import ctypes
import os

import numpy as np

ZSTD_LEVEL = 9                       # compress.py:276,398
MAX_DECODED_BYTES = 1 << 37          # 128 GiB: a 32767-frame 1024x1024x1 v2 stream is 2^37 bytes
KEY_FILE, ENTROPY_FILE, NAMES_FILE = "key_frame.dat", "entropy.dat", "filename.txt"
_ZSTD_c_compressionLevel, _ZSTD_c_contentSizeFlag, _ZSTD_c_nbWorkers, _ZSTD_c_jobSize = 100, 200, 400, 401

_z = None


def _zlib():
    global _z
    if _z is None:
        z = ctypes.CDLL("libzstd.so.1")
        sz, vp, ci = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int
        z.ZSTD_compressBound.restype, z.ZSTD_compressBound.argtypes = sz, [sz]
        z.ZSTD_compress.restype, z.ZSTD_compress.argtypes = sz, [vp, sz, vp, sz, ci]
        z.ZSTD_decompress.restype, z.ZSTD_decompress.argtypes = sz, [vp, sz, vp, sz]
        z.ZSTD_getFrameContentSize.restype, z.ZSTD_getFrameContentSize.argtypes = ctypes.c_ulonglong, [vp, sz]
        z.ZSTD_isError.restype, z.ZSTD_isError.argtypes = ctypes.c_uint, [sz]
        z.ZSTD_createCCtx.restype, z.ZSTD_createCCtx.argtypes = vp, []
        z.ZSTD_freeCCtx.restype, z.ZSTD_freeCCtx.argtypes = sz, [vp]
        z.ZSTD_CCtx_setParameter.restype, z.ZSTD_CCtx_setParameter.argtypes = sz, [vp, ci, ci]
        z.ZSTD_compress2.restype, z.ZSTD_compress2.argtypes = sz, [vp, vp, sz, vp, sz]
        _z = z
    return _z


def zstd_compress(buf, level=ZSTD_LEVEL, workers=0):
    """buf: bytes or a C-contiguous numpy array.  Returns bytes (one zstd frame with content size)."""
    z = _zlib()
    a = np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else \
        np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    n = a.size
    cap = z.ZSTD_compressBound(n)
    dst = np.empty(cap, np.uint8)
    if workers > 0:
        c = z.ZSTD_createCCtx()
        try:
            z.ZSTD_CCtx_setParameter(c, _ZSTD_c_compressionLevel, level)
            z.ZSTD_CCtx_setParameter(c, _ZSTD_c_contentSizeFlag, 1)
            z.ZSTD_CCtx_setParameter(c, _ZSTD_c_nbWorkers, int(workers))
            # libzstd's default job is 4 x the window (32 MB at level 9): a 123 MB stream would keep four workers
            # busy.  Jobs of n / (4 x workers), between 1 and 8 MB, keep them all busy; the compressed size is unchanged
            # to four digits (the jobs still overlap: ZSTD_c_overlapLog is left at its default)
            z.ZSTD_CCtx_setParameter(c, _ZSTD_c_jobSize, int(min(8 << 20, max(1 << 20, n // (4 * int(workers))))))
            r = z.ZSTD_compress2(c, dst.ctypes.data, cap, a.ctypes.data, n)
        finally:
            z.ZSTD_freeCCtx(c)
    else:
        r = z.ZSTD_compress(dst.ctypes.data, cap, a.ctypes.data, n, level)
    if z.ZSTD_isError(r):
        raise RuntimeError("zstd compression failed")
    return dst[:r].tobytes()


def zstd_decompress(data):
    """Returns a numpy uint8 array (one-shot decode sized from the frame header, like python-zstd)."""
    z = _zlib()
    src = np.frombuffer(data, np.uint8)
    size = z.ZSTD_getFrameContentSize(src.ctypes.data, src.size)
    if size >= (1 << 62):
        raise RuntimeError("zstd frame does not carry its content size")
    # the header is untrusted input: refuse absurd sizes instead of attempting the allocation
    cap = int(os.environ.get("TEZIP_MAX_DECODED_BYTES", str(MAX_DECODED_BYTES)))
    if int(size) > cap:
        raise RuntimeError("zstd frame declares %d bytes of content, more than the limit of %d "
                           "(TEZIP_MAX_DECODED_BYTES)" % (int(size), cap))
    try:
        dst = np.empty(max(int(size), 1), np.uint8)
    except MemoryError:
        raise RuntimeError("cannot allocate %d bytes for the decoded stream" % int(size))
    r = z.ZSTD_decompress(dst.ctypes.data, int(size), src.ctypes.data, src.size)
    if z.ZSTD_isError(r) or r != size:
        raise RuntimeError("zstd decompression failed")
    return dst[:r]


def container_level():
    """zstd level of the container files: 9 like the reference (compress.py:276,398) unless TEZIP_ZSTD_LEVEL says
    otherwise -- any level gives a frame the reference's decoder reads; lower levels trade a few per cent of ratio for
    a several times faster container stage (bench.py's `container` record quotes both)."""
    v = os.environ.get("TEZIP_ZSTD_LEVEL", str(ZSTD_LEVEL))
    return ZSTD_LEVEL if v.strip().lower() == "gpu" else int(v)


def gpu_writer():
    """True when TEZIP_ZSTD_LEVEL=gpu: the two frames are written by the CUDA kernels of zstd_frames.py from the device
    copies of key plane and stream (Huffman-coded literal blocks, no match finding: an order-0 ratio at memory speed)
    instead of by libzstd from host copies.  Either way the files are single zstd frames with content size."""
    return os.environ.get("TEZIP_ZSTD_LEVEL", "").strip().lower() == "gpu"


def default_workers():
    """zstd worker threads of the container stage: TEZIP_ZSTD_WORKERS, else every host core (0 = libzstd's
    single-threaded encoder).  Any count gives one frame with content size, which is all the reference's decoder needs."""
    v = os.environ.get("TEZIP_ZSTD_WORKERS")
    return int(v) if v is not None else (os.cpu_count() or 1)


def _write_names(out_dir, names, is_rgb):
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, NAMES_FILE), "w", encoding="UTF-8") as f:     # compress.py:133-136
        f.write("%d\n" % int(is_rgb))
        for nm in names:
            f.write("%s\n" % nm)


def write_container_device(out_dir, names, is_rgb, key_plane, body, tail):
    """write_container with the zstd frames written on the GPU (gpu_writer()): key_plane and body are CUDA tensors
    (u8 + int16, or u16 + int32 for container v2), tail the host array that follows the codes in entropy.dat (table,
    its length, shape, p: codec.pack_payload of an empty body).  -> (key_frame.dat bytes, entropy.dat bytes)."""
    import torch
    from . import zstd_frames
    _write_names(out_dir, names, is_rgb)
    wide = body.dtype == torch.int32
    if wide != (key_plane.dtype == torch.uint16) or np.asarray(tail).dtype != (np.int32 if wide else np.int16):
        raise ValueError("key plane, stream and trailer disagree about the sample width")
    sizes = []
    tail_dev = torch.from_numpy(np.ascontiguousarray(tail)).to(body.device)
    for fn, t in ((KEY_FILE, key_plane),                                                   # compress.py:271-278
                  (ENTROPY_FILE, torch.cat([body.reshape(-1), tail_dev]))):                # compress.py:394-400
        frame = zstd_frames.frame_host(t)          # a view of a pinned buffer: written out before the next frame
        with open(os.path.join(out_dir, fn), "wb") as f:
            f.write(memoryview(frame))
        sizes.append(int(frame.size))
    return sizes[0], sizes[1]


def write_container(out_dir, names, is_rgb, key_plane, payload, workers=None):
    """key_plane: u8 array (any shape); payload: int16 array (entropy.dat before zstd).
    Container v2 (16-bit samples): key_plane u16 and payload int32, both little-endian, same three files.
    The two zstd frames are produced concurrently (libzstd releases the GIL under ctypes)."""
    _write_names(out_dir, names, is_rgb)
    wide = np.asarray(payload).dtype == np.int32
    if wide != (np.asarray(key_plane).dtype == np.uint16):
        raise ValueError("key plane and stream disagree about the sample width")
    level = container_level()
    if workers is None:
        workers = default_workers()
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as pool:
        fk = pool.submit(zstd_compress, np.ascontiguousarray(key_plane, "<u2" if wide else np.uint8), level,
                         max(0, workers // 4))                                                          # compress.py:271-278
        fe = pool.submit(zstd_compress, np.ascontiguousarray(payload, "<i4" if wide else "<i2"), level, workers)  # :394-400
        kb, eb = fk.result(), fe.result()
    with open(os.path.join(out_dir, KEY_FILE), "wb") as f:
        f.write(kb)
    with open(os.path.join(out_dir, ENTROPY_FILE), "wb") as f:
        f.write(eb)
    return len(kb), len(eb)


def read_container(comp_dir):
    """-> (names, is_rgb, key_plane u8 flat, payload int16)  (decompress.py:48-56,87-103); a container-v2 stream
    (recognised by the magic that ends it) -> key_plane u16 flat, payload int32."""
    with open(os.path.join(comp_dir, NAMES_FILE), "r", encoding="UTF-8") as f:
        names = [s.strip() for s in f.readlines()]
    is_rgb = True
    if names and len(names[0]) == 1 and names[0].isdigit():                               # decompress.py:55-56
        is_rgb = bool(int(names.pop(0)))
    with open(os.path.join(comp_dir, KEY_FILE), "rb") as f:
        key_plane = zstd_decompress(f.read())
    with open(os.path.join(comp_dir, ENTROPY_FILE), "rb") as f:
        raw = zstd_decompress(f.read())
    from .codec import is_v2_payload
    if is_v2_payload(raw):
        return names, is_rgb, key_plane.view("<u2"), raw.view("<i4")
    return names, is_rgb, key_plane, raw.view("<i2")


_PINNED_IN = {}


def read_container_device(comp_dir, device):
    """read_container with the two frames decoded on the GPU: -> (names, is_rgb, key plane, payload) as CUDA tensors
    (u8 + int16, or u16 + int32 for container v2), or None when a frame uses parts of the zstd format that the kernels
    of zstd_frames.py do not read (frames written by libzstd: read_container decodes those, like decompress.py:89,98).
    TEZIP_ZSTD_DECODER=cpu turns this path off."""
    if os.environ.get("TEZIP_ZSTD_DECODER", "").strip().lower() == "cpu":
        return None
    import torch
    from . import zstd_frames
    from .codec import V2_MAGIC
    raws = []
    cap = int(os.environ.get("TEZIP_MAX_DECODED_BYTES", str(MAX_DECODED_BYTES)))
    for fn in (KEY_FILE, ENTROPY_FILE):
        path = os.path.join(comp_dir, fn)
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            head = f.read(6)
            if head != b"\x28\xb5\x2f\xfd\xc0\x38":      # not the GPU writer's frame header: libzstd's business
                return None
            buf = _PINNED_IN.get(fn)                       # read into pinned memory (cached buffer): the upload is
            if buf is None or buf.numel() < size:          # asynchronous and overlaps the header walk
                buf = _PINNED_IN[fn] = torch.empty(max(size, 1 << 20), dtype=torch.uint8, pin_memory=True)
            data = buf.numpy()[:size]
            data[:6] = np.frombuffer(head, np.uint8)
            if f.readinto(memoryview(data)[6:]) != size - 6:
                raise RuntimeError("short read of %s" % path)
        got = zstd_frames.decompress_device(data, device, max_bytes=cap)
        if got is None:
            return None
        raws.append(got)
    with open(os.path.join(comp_dir, NAMES_FILE), "r", encoding="UTF-8") as f:
        names = [s.strip() for s in f.readlines()]
    is_rgb = True
    if names and len(names[0]) == 1 and names[0].isdigit():                               # decompress.py:55-56
        is_rgb = bool(int(names.pop(0)))
    key, raw = raws
    n = raw.numel()
    wide = n >= 44 and n % 4 == 0 and int.from_bytes(bytes(raw[-4:].cpu().numpy()), "little") == V2_MAGIC
    if wide:
        if key.numel() % 2:
            raise RuntimeError("key plane of a 16-bit container has an odd number of bytes")
        return names, is_rgb, key.view(torch.uint16), raw.view(torch.int32)
    if n % 2:
        raise RuntimeError("entropy.dat holds an odd number of bytes")
    return names, is_rgb, key, raw.view(torch.int16)
