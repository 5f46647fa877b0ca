"""Drop-in for /root/reference/src/compress.py: same `run(...)` signature (compress.py:93), same container.

Image loading (compress.py:97-131) and the container writer stay on the host; everything in between runs on
the GPU through codec.encode_frames.  Errors print the reference's messages and exit (non-zero, unlike the
reference's bare `exit()`).
"""
import glob
import os
import sys
import time

import numpy as np
import torch

from . import codec, container
from ._lib import TezipError
from .prednet import PredNet


def _die(*msg):
    print(*msg)
    sys.exit(1)


def io_workers():
    """Threads for image decode/encode (PIL releases the GIL inside its codecs); TEZIP_IO_WORKERS overrides."""
    env = os.environ.get("TEZIP_IO_WORKERS")
    return max(1, int(env)) if env else max(1, min(32, os.cpu_count() or 1))


class SequenceLoader:
    """compress.py:97-131 as a stream: the images are decoded IN ORDER by a thread pool into one pinned buffer while
    the caller already works on the frames that have arrived (`wait(a, b)` blocks until frames [a, b) are decoded).
    The reference decodes the whole sequence before its first prediction; here the GPU starts on the first group of
    windows as soon as its frames are there."""

    def __init__(self, DATA_DIR, workers=None, paths=None):
        from concurrent.futures import ThreadPoolExecutor
        from PIL import Image, UnidentifiedImageError
        self.dir = DATA_DIR
        self.errors = (PermissionError, IndexError, UnidentifiedImageError, IsADirectoryError, ValueError)
        file_paths = list(paths) if paths is not None else sorted(glob.glob(os.path.join(DATA_DIR, '*')))
        if len(file_paths) == 0:
            _die("ERROR:", DATA_DIR, "is an empty or non-existent directory")
        self.paths = file_paths
        self.files = [os.path.basename(path) for path in file_paths]
        try:
            first = Image.open(file_paths[0])
            image_mode = first.mode
            wide = image_mode in ('I;16', 'I;16L')
            if image_mode not in ('RGB', 'L') and not wide:
                _die("ERROR: input image is {0}. Only RGB and grayscale are supported.".format(image_mode))
            self.isRGB = isRGB = image_mode == 'RGB'
            w, h = first.size
        except self.errors:
            _die(DATA_DIR, "contains files or folders that are not images.")
        buf = torch.empty((len(file_paths), h, w, 1), dtype=torch.uint16) if wide else \
            torch.empty((len(file_paths), h, w, 3), dtype=torch.uint8)
        if torch.cuda.is_available():
            buf = buf.pin_memory()
        self.tensor = buf
        self.frames = frames = buf.numpy()

        def decode(i):
            img = Image.open(file_paths[i])
            if wide:
                if img.mode not in ('I;16', 'I;16L'):
                    raise ValueError("mixed image modes")
                frames[i, :, :, 0] = np.asarray(img, dtype=np.uint16)
            else:
                frames[i] = np.asarray(img if isRGB else img.convert('RGB'))    # shape mismatch -> ValueError

        self.pool = ThreadPoolExecutor(workers or io_workers())
        self.futures = [self.pool.submit(decode, i) for i in range(len(file_paths))]   # FIFO: decoded in order

    def wait(self, a=0, b=None):
        try:
            for f in self.futures[a:b]:
                f.result()
        except self.errors:
            self.close()
            _die(self.dir, "contains files or folders that are not images.")

    def close(self):
        self.pool.shutdown(wait=False, cancel_futures=True)


def load_images(DATA_DIR, workers=None, paths=None):
    """compress.py:97-131: sorted glob, RGB or L (converted to RGB) -> u8 [nt,H,W,3], basenames, isRGB.
    16-bit grayscale images (PIL mode I;16 -- the reference stops at compress.py:106-110) -> u16 [nt,H,W,1], written
    as a version-2 container.  One pinned allocation for the whole sequence (the reference hstacks frame by frame, O(nt^2)) filled by a thread
    pool, so the array goes to the GPU with one asynchronous copy."""
    loader = SequenceLoader(DATA_DIR, workers, paths)
    loader.wait()
    loader.close()
    return loader.frames, loader.files, loader.isRGB


def save_images(frames, file_names, isRGB, OUTPUT_DIR, workers=None):
    """decompress.py:266-279 with a thread pool.  The reference re-saves every image as RGB (decompress.py:278),
    overwriting its own 'L' conversion; the grayscale branch is honoured here (documented deviation, SURVEY.md
    Appendix B)."""
    from concurrent.futures import ThreadPoolExecutor
    from PIL import Image
    wide = frames.dtype == np.uint16
    print("save as 16-bit gray" if wide else "save as RGB" if isRGB else "save as gray")

    def encode(j):
        if wide:
            Image.fromarray(np.ascontiguousarray(frames[j, :, :, 0])).save(os.path.join(OUTPUT_DIR, file_names[j]))
            return
        img = Image.fromarray(frames[j])
        (img if isRGB else img.convert("L")).save(os.path.join(OUTPUT_DIR, file_names[j]))

    with ThreadPoolExecutor(workers or io_workers()) as pool:
        list(pool.map(encode, range(len(file_names))))


def load_predictor(WEIGHTS_DIR, max_batch, device=0):
    """compress.py:143-173."""
    json_file = os.path.join(WEIGHTS_DIR, 'prednet_model.json')
    if not os.path.exists(json_file):
        _die("ERROR: No such file or directory:", json_file)
    try:
        return PredNet.from_model_dir(WEIGHTS_DIR, max_batch=max_batch, device=device)
    except OSError:
        _die("ERROR: No such file or directory:", os.path.join(WEIGHTS_DIR, 'prednet_weights.hdf5'))


def run_sharded(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, PREPROCESS, WINDOW_SIZE, MODE, BOUND_VALUE, VERBOSE, ENTROPY_RUN,
                zstd_workers=None, THRESHOLD=None, dwp_chains=1):
    """compress.run under torchrun (one process per GPU): every rank decodes and encodes the images of its own shard,
    the ranks exchange the 1-element delta halo and sum the symbol histogram (dist.ShardComm), rank 0 gathers the
    streams and writes the single container.  Static windows (-w): window-aligned shards
    (dist.shard_ranges); the container is byte-identical before zstd to the one a single process writes.  Dynamic
    windows (-t): contiguous frame ranges, every range starting with a forced key frame (container-legal: the
    decoder finds key frames in the key plane), and inside a rank `dwp_chains` sub-ranges run as a batch."""
    import torch.distributed as tdist
    from . import dist as tzdist
    rank, world, local = tzdist.init_from_env()
    file_paths = sorted(glob.glob(os.path.join(DATA_DIR, '*')))
    nt = len(file_paths)
    if nt == 0:
        _die("ERROR:", DATA_DIR, "is an empty or non-existent directory")
    if THRESHOLD is None:
        ranges = tzdist.shard_ranges(nt, PREPROCESS, WINDOW_SIZE, world)
    else:
        ranges = tzdist.shard_ranges(nt, PREPROCESS, max(1, -(-(nt - PREPROCESS) // world)), world)
    a, b = ranges[rank]
    # every rank evaluates the same layout test, so that they all leave together (a rank that exits alone would leave
    # its peers blocked in the first collective)
    if nt < PREPROCESS + 2 or any(rb <= ra for ra, rb in ranges):
        _die("ERROR: fewer windows than GPUs (or fewer than p+2 frames); use fewer processes")
    frames, files, isRGB = load_images(DATA_DIR, paths=file_paths[a:b])
    n_win = max(1, (b - a + WINDOW_SIZE - 1) // WINDOW_SIZE) if THRESHOLD is None else max(1, dwp_chains)
    net = load_predictor(WEIGHTS_DIR, max_batch=min(n_win, 256), device=local)
    dev = net.device
    comm = tzdist.ShardComm(device=dev)
    enc, err = None, None
    try:
        enc = codec.encode_frames(torch.from_numpy(frames).to(dev, non_blocking=True), net,
                                  PREPROCESS if rank == 0 else 0, WINDOW_SIZE, THRESHOLD, MODE, list(BOUND_VALUE),
                                  ENTROPY_RUN, dwp_chains=dwp_chains, comm=comm)
    except TezipError as e:
        err = str(e)
    if not tzdist.all_ok(err is None):      # agree before the stream gather
        _die(err or "ERROR: another rank failed")
    fe = frames[0].size
    sizes = [(rb - ra) * fe for ra, rb in ranges]
    body = tzdist.gather_varlen(enc.body, sizes)
    keyp = tzdist.gather_varlen(enc.key_plane.reshape(-1), sizes)
    names = [None] * world
    tdist.all_gather_object(names, files)
    if rank == 0:
        H, W, C = frames.shape[1:]
        if container.gpu_writer():      # TEZIP_ZSTD_LEVEL=gpu: the frames are written from the device copies
            tail = codec.pack_payload(body[:0].cpu().numpy(), enc.table, (1, nt, H, W, C), PREPROCESS)
            kb, eb = container.write_container_device(OUTPUT_DIR, [f for part in names for f in part], isRGB,
                                                      keyp, body, tail)
        else:
            payload = codec.pack_payload(body.cpu().numpy(), enc.table, (1, nt, H, W, C), PREPROCESS)
            kb, eb = container.write_container(OUTPUT_DIR, [f for part in names for f in part], isRGB,
                                               keyp.cpu().numpy(), payload, workers=zstd_workers)
        if VERBOSE:
            print("ranks:", world, "ratio:", nt * fe / float(kb + eb))
    tdist.barrier()
    net.close()


def run(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, PREPROCESS, WINDOW_SIZE, THRESHOLD, MODE, BOUND_VALUE, GPU_FLAG, VERBOSE,
        ENTROPY_RUN, dwp_chains=1, zstd_workers=None):
    if not GPU_FLAG:
        _die("ERROR: tezip_b200 has no CPU path; a B200 (sm_100) GPU is required.")
    from . import dist as tzdist
    if tzdist.launched_by_torchrun():
        if not os.path.exists(OUTPUT_DIR):
            os.makedirs(OUTPUT_DIR, exist_ok=True)
        return run_sharded(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, PREPROCESS, WINDOW_SIZE, MODE, BOUND_VALUE, VERBOSE,
                           ENTROPY_RUN, zstd_workers, THRESHOLD, dwp_chains)
    if not os.path.exists(OUTPUT_DIR):
        os.mkdir(OUTPUT_DIR)
    loader = SequenceLoader(DATA_DIR)          # decoding starts now, in order, on the I/O threads
    frames, files, isRGB = loader.frames, loader.files, loader.isRGB
    nt = frames.shape[0]
    n_win = max(1, (nt - PREPROCESS + (WINDOW_SIZE or nt) - 1) // (WINDOW_SIZE or nt)) if THRESHOLD is None \
        else max(1, dwp_chains)
    # static windows: groups of STREAM_WINDOWS windows are predicted as their frames arrive (the weights load and the
    # first groups' kernels overlap the decoding of the rest); dynamic windows scan the whole sequence
    stream = THRESHOLD is None and nt > 0
    group = int(os.environ.get("TEZIP_STREAM_WINDOWS", "32"))
    net = load_predictor(WEIGHTS_DIR, max_batch=min(max(n_win, 1), group if stream else 256))
    try:
        t0 = time.time()
        dev = net.device
        if stream:
            frames_dev = torch.empty(loader.tensor.shape, dtype=loader.tensor.dtype, device=dev)

            def arrive(a, b):          # frames [a, b) are needed on the device now
                loader.wait(a, b)
                frames_dev[a:b].copy_(loader.tensor[a:b], non_blocking=True)
            enc = codec.encode_frames(frames_dev, net, PREPROCESS, WINDOW_SIZE, None, MODE, list(BOUND_VALUE),
                                      ENTROPY_RUN, arrive=arrive)
        else:
            loader.wait()
            enc = codec.encode_frames(loader.tensor.to(dev, non_blocking=True), net, PREPROCESS, WINDOW_SIZE,
                                      THRESHOLD, MODE, list(BOUND_VALUE), ENTROPY_RUN, dwp_chains=dwp_chains)
        loader.close()
        on_gpu = container.gpu_writer()         # TEZIP_ZSTD_LEVEL=gpu: the frames are written from the device copies
        if not on_gpu:
            payload = enc.payload()
            key_plane = enc.key_plane.cpu().numpy()
        torch.cuda.synchronize(dev)
        if VERBOSE:
            print("gpu_encode:{0}".format(time.time() - t0) + "[sec]")
        t0 = time.time()
        if on_gpu:
            tail = codec.pack_payload(enc.body[:0].cpu().numpy(), enc.table, enc.shape, enc.p)
            kb, eb = container.write_container_device(OUTPUT_DIR, files, isRGB, enc.key_plane, enc.body, tail)
    except TezipError as e:
        _die(str(e))
    if not on_gpu:
        kb, eb = container.write_container(OUTPUT_DIR, files, isRGB, key_plane, payload, workers=zstd_workers)
    if VERBOSE:
        print("zstd+write:{0}".format(time.time() - t0) + "[sec]")
        print("key frames:", len(enc.keys), "ratio:", frames.nbytes / float(kb + eb))
    net.close()
