"""Drop-in for /root/reference/src/decompress.py: same `run(...)` signature (decompress.py:39), same container."""
import os
import sys
import time

import numpy as np
import torch

from . import codec, container
from ._lib import TezipError
from .compress import load_predictor, save_images


def _die(*msg):
    print(*msg)
    sys.exit(1)


def run_sharded(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, VERBOSE):
    """decompress.run under torchrun: rank 0 reads the container and finds the key frames, the windows are dealt out
    as contiguous key-aligned frame ranges, every rank decodes its range (x restarts at 0 at its first key frame,
    SURVEY.md A18) and writes its own images."""
    import torch.distributed as tdist
    from . import dist as tzdist, ops
    rank, world, local = tzdist.init_from_env()
    dev = torch.device("cuda", local)
    meta = [None]
    key_full = body_full = None
    fail = None
    if rank == 0:
        try:
            file_names, isRGB, key_plane, payload = container.read_container(DATA_DIR)
            body, table, shape, p = codec.parse_payload(payload)
            if len(file_names) != shape[1]:
                fail = "ERROR：The lengths of filename.txt and images do not match. number of images %d" % shape[1]
        except (TezipError, RuntimeError, OSError) as e:
            fail = str(e)
    status = [fail]
    tdist.broadcast_object_list(status, src=0)      # every rank leaves together if rank 0 could not read the container
    if status[0] is not None:
        _die(status[0])
    if rank == 0:
        _one, nt, H, W, C = shape
        key_full = torch.from_numpy(np.ascontiguousarray(key_plane)).to(dev)
        body_full = torch.from_numpy(np.ascontiguousarray(body)).to(dev)
        nz = ops.frames_nonzero(key_full.view(nt, H, W, C)).cpu().numpy()
        keys = [int(i) for i in np.nonzero(nz)[0]]
        ranges = tzdist.key_aligned_ranges(keys, nt, p, world)
        meta = [(file_names, isRGB, None if table is None else np.asarray(table), shape, p, ranges,
                 body.dtype == np.int32)]
    tdist.broadcast_object_list(meta, src=0)
    file_names, isRGB, table, shape, p, ranges, _wide = meta[0]
    _one, nt, H, W, C = shape
    fe = H * W * C
    sizes = [(b - a) * fe for a, b in ranges]
    a, b = ranges[rank]
    wide = meta[0][6]
    key_part = tzdist.scatter_varlen(key_full, sizes, torch.uint16 if wide else torch.uint8, dev)
    body_part = tzdist.scatter_varlen(body_full, sizes, torch.int32 if wide else torch.int16, dev)
    if b > a:
        n_keys_guess = max(1, (b - a) // 4)
        net = load_predictor(WEIGHTS_DIR, max_batch=min(n_keys_guess, 256), device=local)
        err = None
        try:
            out, _plan = codec.decode_arrays(key_part.contiguous(), body_part.contiguous(), table, (1, b - a, H, W, C),
                                             p if rank == 0 else 0, net, first_mode=0 if rank == 0 else 1, first_x=0)
        except TezipError as e:
            err = str(e)
        if err is None:
            save_images(out.cpu().numpy(), file_names[a:b], isRGB, OUTPUT_DIR)
        net.close()
    else:
        err = None
    if not tzdist.all_ok(err is None):
        _die(err or "ERROR: another rank failed")


def run(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, GPU_FLAG, VERBOSE):
    if not GPU_FLAG:
        _die("ERROR: tezip_b200 has no CPU path; a B200 (sm_100) GPU is required.")
    from . import dist as tzdist
    if tzdist.launched_by_torchrun():
        os.makedirs(OUTPUT_DIR, exist_ok=True)
        for fn in (container.NAMES_FILE, container.KEY_FILE, container.ENTROPY_FILE):
            if not os.path.exists(os.path.join(DATA_DIR, fn)):
                _die("ERROR: No such file or directory:", os.path.join(DATA_DIR, fn))
        return run_sharded(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, VERBOSE)
    if not os.path.exists(OUTPUT_DIR):
        os.mkdir(OUTPUT_DIR)
    for fn in (container.NAMES_FILE, container.KEY_FILE, container.ENTROPY_FILE):
        if not os.path.exists(os.path.join(DATA_DIR, fn)):
            _die("ERROR: No such file or directory:", os.path.join(DATA_DIR, fn))     # decompress.py:51-53,90-101
    # frames written by the GPU writer (TEZIP_ZSTD_LEVEL=gpu) are decoded on the device and never exist on the host;
    # anything else goes through libzstd like decompress.py:89,98
    on_dev = container.read_container_device(DATA_DIR, torch.device("cuda", torch.cuda.current_device())) \
        if torch.cuda.is_available() else None
    if on_dev is not None:
        file_names, isRGB, key_plane, payload_dev = on_dev
        tail_len = min(payload_dev.numel(), 1 << 19)         # trailer + table (at most 262144 + 11 entries)
        payload = payload_dev[payload_dev.numel() - tail_len:].cpu().numpy()
    else:
        file_names, isRGB, key_plane, payload = container.read_container(DATA_DIR)
    try:
        try:
            body, table, shape, p = codec.parse_payload(payload)
        except TezipError:
            if on_dev is None or tail_len == payload_dev.numel():
                raise
            payload = payload_dev.cpu().numpy()              # (a table longer than the tail that was fetched)
            tail_len = payload.size
            body, table, shape, p = codec.parse_payload(payload)
        if len(file_names) != shape[1]:                                                # decompress.py:260-264
            print("ERROR：The lengths of filename.txt and images do not match.")
            print("filename.txt：", len(file_names))
            _die("number of images", shape[1])
        n_keys_guess = max(1, shape[1] // 4)
        net = load_predictor(WEIGHTS_DIR, max_batch=min(n_keys_guess, 256))
        dev = net.device
        t0 = time.time()
        if on_dev is not None:
            n_body = payload_dev.numel() - (tail_len - body.size)
            out, _plan = codec.decode_arrays(key_plane.to(dev), payload_dev[:n_body].to(dev), table, shape, p, net)
        else:
            out, _plan = codec.decode_arrays(torch.from_numpy(np.ascontiguousarray(key_plane)).to(dev),
                                             torch.from_numpy(np.ascontiguousarray(body)).to(dev), table, shape, p, net)
        frames = out.cpu().numpy()
        if VERBOSE:
            print("gpu_decode:{0}".format(time.time() - t0) + "[sec]")
    except TezipError as e:
        _die(str(e))
    save_images(frames, file_names, isRGB, OUTPUT_DIR)
    net.close()
