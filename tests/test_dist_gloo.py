"""World-size-2 gloo test of the sharding protocol (tezip_b200/dist.py): with the halo all-gather and the histogram
all-reduce, per-shard delta + rank-map streams concatenate to exactly the single-process stream.  The per-shard
kernels are emulated with the oracle's functions (the CUDA kernels need a GPU; their has_prev/prev_x arguments are
covered by tests/test_gpu_dist.py)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, x_all, ranges, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tezip_b200 import ops
    from tezip_b200.dist import ShardComm
    comm = ShardComm(None, None)
    a, b = ranges[rank]
    x = x_all[a:b].astype(np.int16)
    has_prev, prev_x = comm.exchange_last_x(int(x[-1]))
    assert has_prev == (rank > 0)
    y = np.empty_like(x)
    y[1:] = x[:-1] - x[1:]
    y[0] = (np.int16(prev_x) - x[0]) if has_prev else x[0]
    hist = torch.from_numpy(np.bincount((1600 - y).astype(np.int64), minlength=4096).astype(np.int64))
    comm.reduce_hist(hist)
    table = ops.build_table(hist.numpy())
    body = ops.encode_lut(table)[(1600 - y).astype(np.int16)]
    offs, sizes = comm.stream_offsets(body.size)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), body=body, table=table, off=offs[rank], sizes=sizes)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stream_equals_single_process(tmp_path):
    from oracle import codec_oracle as co
    rng = np.random.default_rng(5)
    x_all = rng.integers(-20, 21, 6000).astype(np.int16)
    world = 2
    ranges = [(0, 2400), (2400, 6000)]
    mp.spawn(_worker, args=(world, _free_port(), x_all, ranges, str(tmp_path)), nprocs=world, join=True)
    y = co.delta_encode(x_all)
    s = (1600 - y).astype(np.int16)
    table = co.build_table(s)
    ref = co.replacing_encode(s, table)
    parts = [np.load(str(tmp_path / ("r%d.npz" % r))) for r in range(world)]
    assert all(np.array_equal(p["table"], table) for p in parts)
    assert [int(p["off"]) for p in parts] == [0, 2400]
    assert np.array_equal(np.concatenate([p["body"] for p in parts]), ref)
