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
        """x_last: device int32[1] (this shard's last residual).  -> (has_prev, device int32[1] holding the previous
        shard's last residual, or None on rank 0).  No host synchronisation: the gathered array stays on the device
        and the delta kernels read it through a pointer (tz_delta_hist's prev_x)."""
        if not torch.is_tensor(x_last):
            x_last = torch.tensor([int(x_last)], dtype=torch.int32, device=self.device if self.backend == "nccl" else "cpu")
        mine = self._stage(x_last.reshape(1).to(torch.int32))
        allx = torch.zeros(self.world, dtype=torch.int32, device=mine.device)
        dist.all_gather_into_tensor(allx, mine, group=self.group) if self.backend == "nccl" else \
            allx.copy_(torch.cat(self._gather_list(mine)))
        if self.rank == 0:
            return False, None
        prev = allx[self.rank - 1:self.rank]
        return True, (prev if prev.device == x_last.device else prev.to(x_last.device))

    def _gather_list(self, mine):
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return out

    def reduce_hist(self, hist):
        t = self._stage(hist)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        if t is not hist:
            hist.copy_(t)

    def stream_offsets_async(self, n_local):
        """All-gather of the shards' stream lengths, queued without a host synchronisation.  Returns a callable that
        yields (offsets, sizes) -- offsets = exclusive prefix sum = each shard's position in entropy.dat -- and only
        then waits for the collective."""
        dev = self.device if self.backend == "nccl" else "cpu"
        mine = torch.full((1,), int(n_local), dtype=torch.int64, device=dev)
        alln = torch.zeros(self.world, dtype=torch.int64, device=dev)
        if self.backend == "nccl":
            dist.all_gather_into_tensor(alln, mine, group=self.group)
        else:
            alln.copy_(torch.cat(self._gather_list(mine)))

        def result():
            sizes = alln.cpu().numpy().astype(np.int64)
            return np.concatenate([[0], np.cumsum(sizes)[:-1]]), sizes
        return result

    def stream_offsets(self, n_local):
        return self.stream_offsets_async(n_local)()


# ------------------------------------------------------------------------------------------------ file drivers
def launched_by_torchrun():
    """True under `torchrun` / `python -m torch.distributed.run` with more than one rank."""
    import os
    return int(os.environ.get("WORLD_SIZE", "1")) > 1


def init_from_env():
    """One process per GPU (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the launcher).  -> (rank, world, device)."""
    import os
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, world, local


def gather_varlen(t, sizes, dst=0, group=None):
    """Concatenation, on rank `dst`, of every rank's 1-D tensor `t` (rank r holds sizes[r] elements); None elsewhere.
    Shards are padded to the largest size so that one gather moves them."""
    if t.dtype not in (torch.uint8, torch.int32, torch.int64, torch.float32, torch.float16):   # e.g. int16: not an NCCL type
        k = t.element_size()
        out = gather_varlen(t.reshape(-1).view(torch.uint8), [int(v) * k for v in sizes], dst, group)
        return None if out is None else out.view(t.dtype)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = int(max(sizes))
    pad = torch.zeros(n, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t.reshape(-1)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:int(sizes[r])] for r, b in enumerate(bufs)])


def scatter_varlen(full, sizes, dtype, device, src=0, group=None):
    """Inverse of gather_varlen: rank r receives elements [sum(sizes[:r]), sum(sizes[:r+1])) of `full` (given on src)."""
    if dtype not in (torch.uint8, torch.int32, torch.int64, torch.float32, torch.float16):
        k = torch.empty(0, dtype=dtype).element_size()
        full8 = None if full is None else full.reshape(-1).view(torch.uint8)
        return scatter_varlen(full8, [int(v) * k for v in sizes], torch.uint8, device, src, group).view(dtype)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = int(max(sizes))
    out = torch.empty(n, dtype=dtype, device=device)
    parts = None
    if rank == src:
        offs = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))])
        parts = []
        for r in range(world):
            buf = torch.zeros(n, dtype=dtype, device=device)
            buf[:int(sizes[r])] = full[int(offs[r]):int(offs[r + 1])]
            parts.append(buf)
    dist.scatter(out, parts, src=src, group=group)
    return out[:int(sizes[rank])]


def key_aligned_ranges(keys, nt, p, world):
    """Contiguous frame ranges [a, b) per rank for DECODING: every range but the first starts at a key frame >= p
    (x restarts at 0 there), windows are dealt out evenly.  keys: sorted key-frame indices."""
    starts = [k for k in keys if k >= p]
    n_win = len(starts)
    out = []
    for r in range(world):
        # ceil split: with fewer windows than ranks the LEADING ranks get one window each (rank 0 must own a real
        # window: a range of warm-up frames only cannot be decoded), the trailing ranks get empty ranges
        w0, w1 = -(-n_win * r // world), -(-n_win * (r + 1) // world)
        a = 0 if r == 0 else (starts[w0] if w0 < n_win else nt)
        b = nt if r == world - 1 else (starts[w1] if w1 < n_win else nt)
        out.append((min(a, nt), min(b, nt)))
    return out


def all_ok(ok, group=None):
    """True on every rank only if `ok` is true on every rank (ranks must fail together: one that exits alone leaves
    its peers blocked in the next collective)."""
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())
