"""Multi-GPU: windows are independent, so a sequence is sharded by whole windows across the ranks of one box
(one process per GPU, torch.distributed).  The data path has no collective; NCCL carries only
  * all-gather of each shard's last residual (the 1-element halo of the 1-D delta, compress.py:75),
  * all-reduce of the 4096-bin symbol histogram (every rank must build the same table, compress.py:352-361),
  * all-gather of stream lengths -> exclusive prefix sum = each shard's offset in entropy.dat.
Decoding needs no halo: x restarts at 0 on every key frame.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(nt, p, window, world):
    """Contiguous frame ranges [a, b) per rank, every range starting on a key frame of the static-window
    schedule (p + k*window); rank 0 also owns the warm-up frames."""
    n_win = max(1, -(-(nt - p) // window))
    out = []
    for r in range(world):
        w0, w1 = n_win * r // world, n_win * (r + 1) // world
        a = 0 if r == 0 else p + w0 * window
        b = nt if r == world - 1 else p + w1 * window
        out.append((min(a, nt), min(b, nt)))
    return out


class ShardComm:
    """Collectives of one sharded encode.  Works with the nccl backend (device tensors) and with gloo
    (CPU tests: tensors are staged through the host)."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device

    def _stage(self, t):
        return t if self.backend == "nccl" else t.cpu()

    def exchange_last_x(self, x_last):
        dev = self.device if self.backend == "nccl" else "cpu"
        mine = torch.tensor([int(x_last)], dtype=torch.int32, device=dev)
        allx = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allx, mine, group=self.group)
        if self.rank == 0:
            return False, 0
        return True, int(allx[self.rank - 1].item())

    def reduce_hist(self, hist):
        t = self._stage(hist)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        if t is not hist:
            hist.copy_(t)

    def stream_offsets(self, n_local):
        dev = self.device if self.backend == "nccl" else "cpu"
        mine = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
        alln = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(alln, mine, group=self.group)
        sizes = np.array([int(t.item()) for t in alln], np.int64)
        return np.concatenate([[0], np.cumsum(sizes)[:-1]]), sizes
