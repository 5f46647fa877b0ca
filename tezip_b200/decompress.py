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


def run(WEIGHTS_DIR, DATA_DIR, OUTPUT_DIR, GPU_FLAG, VERBOSE):
    if not GPU_FLAG:
        _die("ERROR: tezip_b200 has no CPU path; a B200 (sm_100) GPU is required.")
    if not os.path.exists(OUTPUT_DIR):
        os.mkdir(OUTPUT_DIR)
    for fn in (container.NAMES_FILE, container.KEY_FILE, container.ENTROPY_FILE):
        if not os.path.exists(os.path.join(DATA_DIR, fn)):
            _die("ERROR: No such file or directory:", os.path.join(DATA_DIR, fn))     # decompress.py:51-53,90-101
    file_names, isRGB, key_plane, payload = container.read_container(DATA_DIR)
    try:
        body, table, shape, p = codec.parse_payload(payload)
        if len(file_names) != shape[1]:                                                # decompress.py:260-264
            print("ERROR：The lengths of filename.txt and images do not match.")
            print("filename.txt：", len(file_names))
            _die("number of images", shape[1])
        n_keys_guess = max(1, shape[1] // 4)
        net = load_predictor(WEIGHTS_DIR, max_batch=min(n_keys_guess, 256))
        dev = net.device
        t0 = time.time()
        out, _plan = codec.decode_arrays(torch.from_numpy(np.ascontiguousarray(key_plane)).to(dev),
                                         torch.from_numpy(np.ascontiguousarray(body)).to(dev), table, shape, p, net)
        frames = out.cpu().numpy()
        if VERBOSE:
            print("gpu_decode:{0}".format(time.time() - t0) + "[sec]")
    except TezipError as e:
        _die(str(e))
    save_images(frames, file_names, isRGB, OUTPUT_DIR)
    net.close()
