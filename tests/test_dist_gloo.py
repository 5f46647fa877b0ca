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
    y[0] = (np.int16(int(prev_x)) - x[0]) if has_prev else x[0]   # prev_x: a one-element int32 tensor
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


def _worker_gather(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tezip_b200.dist import gather_varlen, scatter_varlen
    sizes = [5, 9]
    # int16 goes through the byte view (NCCL has no int16), uint8 directly
    mine16 = torch.arange(sizes[rank], dtype=torch.int16) + 100 * (rank + 1) - 300
    mine8 = (torch.arange(sizes[rank]) + 10 * rank).to(torch.uint8)
    g16, g8 = gather_varlen(mine16, sizes), gather_varlen(mine8, sizes)
    if rank == 0:
        assert g16.dtype == torch.int16 and g16.tolist() == [-200 + i for i in range(5)] + [-100 + i for i in range(9)]
        assert g8.tolist() == list(range(5)) + [10 + i for i in range(9)]
    else:
        assert g16 is None and g8 is None
    back16 = scatter_varlen(g16, sizes, torch.int16, "cpu")
    back8 = scatter_varlen(g8, sizes, torch.uint8, "cpu")
    assert torch.equal(back16, mine16) and torch.equal(back8, mine8)
    open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    dist.barrier()
    dist.destroy_process_group()


def test_gather_scatter_varlen_world2(tmp_path):
    """The collectives of the multi-GPU file drivers (compress.run_sharded / decompress.run_sharded)."""
    world = 2
    mp.spawn(_worker_gather, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(str(tmp_path / ("ok%d" % r))) for r in range(world))


def test_key_aligned_ranges():
    from tezip_b200.dist import key_aligned_ranges
    keys = [0, 1, 2, 7, 12, 17, 22]          # p = 2: warm-up frames 0, 1; real windows start at 2, 7, 12, 17, 22
    for world in (1, 2, 3, 5, 8):
        rs = key_aligned_ranges(keys, 25, 2, world)
        assert rs[0][0] == 0 and rs[-1][1] == 25 and len(rs) == world
        for (a, b), (c, d) in zip(rs, rs[1:]):
            assert b == c and a <= b
        for r, (a, b) in enumerate(rs):
            if r > 0 and b > a:
                assert a in keys and a >= 2
