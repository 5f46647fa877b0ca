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


def load_images(DATA_DIR):
    """compress.py:97-131: sorted glob, RGB or L (converted to RGB) -> u8 [nt,H,W,3], basenames, isRGB."""
    from PIL import Image, UnidentifiedImageError
    file_paths = sorted(glob.glob(os.path.join(DATA_DIR, '*')))
    if len(file_paths) == 0:
        _die("ERROR:", DATA_DIR, "is an empty or non-existent directory")
    try:
        first = Image.open(file_paths[0])
        image_mode = first.mode
        if image_mode not in ('RGB', 'L'):
            _die("ERROR: input image is {0}. Only RGB and grayscale are supported.".format(image_mode))
        isRGB = image_mode == 'RGB'
        w, h = first.size
        frames = np.empty((len(file_paths), h, w, 3), np.uint8)       # one allocation instead of nt hstacks
        files = []
        for i, path in enumerate(file_paths):
            img = Image.open(path)
            frames[i] = np.array(img if isRGB else img.convert('RGB'))
            files.append(os.path.basename(path))
    except (PermissionError, IndexError, UnidentifiedImageError, IsADirectoryError, ValueError):
        _die(DATA_DIR, "contains files or folders that are not images.")
    return frames, files, isRGB


def load_predictor(WEIGHTS_DIR, max_batch, device=0):
    """compress.py:143-173."""
    json_file = os.path.join(WEIGHTS_DIR, 'prednet_model.json')
    if not os.path.exists(json_file):
        _die("ERROR: No such file or directory:", json_file)
    try:
        return PredNet.from_model_dir(WEIGHTS_DIR, max_batch=max_batch, device=device)
    except OSError:
        _die("ERROR: No such file or directory:", os.path.join(WEIGHTS_DIR, 'prednet_weights.hdf5'))


def run(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, PREPROCESS, WINDOW_SIZE, THRESHOLD, MODE, BOUND_VALUE, GPU_FLAG, VERBOSE,
        ENTROPY_RUN, dwp_chains=1, zstd_workers=0):
    if not GPU_FLAG:
        _die("ERROR: tezip_b200 has no CPU path; a B200 (sm_100) GPU is required.")
    if not os.path.exists(OUTPUT_DIR):
        os.mkdir(OUTPUT_DIR)
    frames, files, isRGB = load_images(DATA_DIR)
    nt = frames.shape[0]
    n_win = max(1, (nt - PREPROCESS + (WINDOW_SIZE or nt) - 1) // (WINDOW_SIZE or nt)) if THRESHOLD is None \
        else max(1, dwp_chains)
    net = load_predictor(WEIGHTS_DIR, max_batch=min(max(n_win, 1), 256))
    try:
        t0 = time.time()
        dev = net.device
        enc = codec.encode_frames(torch.from_numpy(frames).to(dev), net, PREPROCESS, WINDOW_SIZE, THRESHOLD, MODE,
                                  list(BOUND_VALUE), ENTROPY_RUN, dwp_chains=dwp_chains)
        payload = enc.payload()
        key_plane = enc.key_plane.cpu().numpy()
        torch.cuda.synchronize(dev)
        if VERBOSE:
            print("gpu_encode:{0}".format(time.time() - t0) + "[sec]")
    except TezipError as e:
        _die(str(e))
    t0 = time.time()
    kb, eb = container.write_container(OUTPUT_DIR, files, isRGB, key_plane, payload, workers=zstd_workers)
    if VERBOSE:
        print("zstd+write:{0}".format(time.time() - t0) + "[sec]")
        print("key frames:", len(enc.keys), "ratio:", frames.size / float(kb + eb))
    net.close()
