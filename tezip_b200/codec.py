"""Array-level compress / decompress: the window scheduler (compress.py:188-268, decompress.py:138-189) driving
batched PredNet steps, and the codec kernels (compress.py:293-395, decompress.py:201-256) through the C ABI.

Everything between "frames on the device" and "int16 stream + key plane on the device" happens here; file I/O,
zstd and the CLI sit in compress.py / decompress.py / container.py.

Scheduling.  Windows are independent, and inside a window frame k is predicted from the prediction of frame
k-1 (from the key frame for k = 1).  So all windows advance in lock step: step k runs ONE batched
`PredNet.next` over every window that is longer than k.  Windows are ordered by length (descending, stable) so
that the live set is always a prefix and each step reads / writes one contiguous block of the prediction pool.
"""
import ctypes
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, ops
from ._lib import TezipError, TZ_HIST_BINS, TZ_WIDE_BINS, check

# Container v2 (16-bit samples): the int32 trailer ends with bits, version and this magic.  Read as the
# reference's int16 trailer the last two fields would be C = 0x3230 and p = 0x5A54 -- impossible values (the
# reference always writes C = 3, compress.py:116-121), so no 8-bit container can be mistaken for a v2 one.
V2_MAGIC = 0x5A543230
V2_VERSION = 2


# ------------------------------------------------------------------------------------------------ schedules
def padding_size(num):
    """data_utils.py:103-107."""
    return num if num % 8 == 0 else (int(num / 8) + 1) * 8


def swp_keys(nt, p, window, shard=False):
    """Static-window key frames (compress.py:189-190,217-220,249-263): warm-up frames 0..p-1, then
    p, p+W, p+2W, ...  shard=True: the frames are one rank's part of a longer sequence -- a trailing one-frame
    window (a lone key frame) is then a legal shard."""
    if nt < p + (1 if shard else 2):
        raise TezipError("need at least p+2 frames (the reference crashes otherwise, compress.py:267)")
    if window < 1:
        raise TezipError("window size must be >= 1")
    return list(range(p)) + list(range(p, nt, window))


@dataclass
class Plan:
    """Which prediction-pool slot every frame uses, and the batched steps that fill the pool."""
    nt: int
    p: int
    keys: list                       # key-frame indices as the decoder will see them (warm-up frames included)
    windows: list                    # (first_frame, n_frames) of the real windows (first_frame >= p)
    pred_slot: np.ndarray            # int32 [nt]: pool slot, or -1 where the frame is its own "prediction"
    apply_eb: np.ndarray             # uint8 [nt]: 1 where error_bound runs (compress.py:315-319)
    steps: list = field(default_factory=list)   # [(key_idx int32[B] | None, slot0, B)] for k = 1, 2, ...
    n_slots: int = 1
    dev: dict = field(default_factory=dict)     # (device, name) -> device copy of a table (cached_plan)


def plan_from_keys(nt, p, keys):
    """Builds the lock-step plan from a key-frame list (used by SWP compress and by every decode)."""
    keys = sorted(int(k) for k in keys)
    real = [k for k in keys if k >= p]
    if not real or real[0] != p:
        raise TezipError("frame %d must be a key frame" % p)
    bounds = real + [nt]
    windows = [(bounds[i], bounds[i + 1] - bounds[i]) for i in range(len(real))]
    pred_slot = np.full(nt, -1, np.int32)
    apply_eb = np.zeros(nt, np.uint8)
    if p > 1:
        pred_slot[1:p] = 0                      # warm-up frames 1..p-1 are coded against P0 (compress.py:197-206)
    order = sorted(range(len(windows)), key=lambda i: -windows[i][1])   # stable: ties keep stream order
    firsts = np.array([windows[i][0] for i in order], np.int32)
    lens = np.array([windows[i][1] for i in order], np.int64)
    steps, slot = [], 1
    k = 1
    while True:
        B = int(np.count_nonzero(lens > k))
        if B == 0:
            break
        steps.append((firsts[:B].copy() if k == 1 else None, slot, B))
        pred_slot[firsts[:B] + k] = slot + np.arange(B, dtype=np.int32)
        apply_eb[firsts[:B] + k] = 1
        slot += B
        k += 1
    return Plan(nt, p, keys, windows, pred_slot, apply_eb, steps, slot)


_PLAN_CACHE = {}


def cached_plan(nt, p, keys):
    """plan_from_keys() memoised on (nt, p, keys): a streaming compressor calls with the same schedule again and
    again, and building it (Python loops over the windows, ~0.1 ms per 1000 frames) happens while the GPU idles at
    the start of every call.  The Plan also carries the device copies of its tables (plan.dev)."""
    key = (int(nt), int(p), tuple(int(k) for k in keys))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= 32:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = _PLAN_CACHE[key] = plan_from_keys(nt, p, keys)
    return plan


def _plan_dev(plan, device, name, make):
    """Device copy of one of the plan's (read-only) tables, uploaded once per device."""
    k = (str(device), name)
    t = plan.dev.get(k)
    if t is None:
        t = plan.dev[k] = make().to(device)
    return t


def run_plan(net, frames, plan, pool, arrive=None):
    """Fills pool[1:] with the predictions of every non-key frame; pool[0] = P0.
    arrive(a, b): optional, called before a group of windows is started with the frame range [a, b) that must be on
    the device by then (streaming loader, compress.run): ranges are contiguous and cover [0, nt) in order.

    Windows are independent, so they are taken in groups of at most net.max_batch (in the plan's length-sorted
    order): a group runs all of its lock-steps back to back, the first from the key frames, every later one as a
    chained step on the previous prediction (its live windows are a prefix of the previous step's)."""
    net.p0(out=pool[0])
    if not plan.steps:
        if arrive is not None:
            arrive(0, plan.nt)
        return
    mb = net.max_batch
    key_idx_all, _slot0, B1 = plan.steps[0]
    done = 0
    for g0 in range(0, B1, mb):
        g1 = min(g0 + mb, B1)
        if arrive is not None:
            # the group's windows in stream order end at the largest (first + length); everything before that is
            # requested now (windows are sorted by length, so for equal windows this is simply the next range)
            mine = set(int(v) for v in key_idx_all[g0:g1])
            end = plan.nt if g1 == B1 else max(f + n for f, n in plan.windows if f in mine)
            if end > done:
                arrive(done, end)
                done = end
        idx = _plan_dev(plan, frames.device, ("key_idx", g0, g1),
                        lambda: torch.from_numpy(np.ascontiguousarray(key_idx_all[g0:g1])))
        x = ops.pad_normalize(frames, idx, net.Hp, net.Wp)                  # compress.py:219 / decompress.py:161
        for k, (_kidx, slot0, B) in enumerate(plan.steps):
            nb = min(B, g1) - g0                                            # windows of this group longer than k + 1
            if nb <= 0:
                break
            out = pool[slot0 + g0:slot0 + g0 + nb]
            if k == 0:
                net.next(x[:nb], out=out)                                   # compress.py:222-229
            else:
                net.next_chained(out)
    if arrive is not None and done < plan.nt:
        arrive(done, plan.nt)


def run_dwp(net, frames, p, threshold, pool, n_chains=1, window=None, shard=False):
    """Dynamic-window scheduler (compress.py:214-266 with THRESHOLD): a chain closes its window at frame idx
    when the mean squared error over the padded frames key+1..idx exceeds `threshold`; idx then becomes the
    next key and its prediction is dropped.  n_chains > 1 splits [p, nt) into contiguous sub-ranges that run
    as a batch, each starting with a forced key frame (container-legal; key placement then differs from the
    sequential reference at the sub-range starts); n_chains = 1 is the reference's sequential scan.

    The whole scan runs on the device without a host round trip per step: every chain advances one frame per step
    whether or not its window closes, so the batch size of every step is known up front; which input a chain reads
    (key frame or its previous prediction), the close decision and the frame -> slot table are device state
    (tz_dwp_gather / tz_window_sse / tz_dwp_update).  Returns DEVICE tensors (is_key u8[nt], pred_slot int32[nt],
    apply u8[nt]) and the number of pool slots used."""
    nt = frames.shape[0]
    if nt < p + (1 if shard else 2):     # (one rank's part of a longer sequence may be a lone key frame)
        raise TezipError("need at least p+2 frames (the reference crashes otherwise, compress.py:267)")
    dev = frames.device
    Hp, Wp, C = net.frame_shape()
    denom = float(Hp * Wp * C)
    n_chains = max(1, min(int(n_chains), nt - p))
    edges = [p + (nt - p) * c // n_chains for c in range(n_chains + 1)]
    spans = [(edges[c], edges[c + 1]) for c in range(n_chains) if edges[c + 1] > edges[c]]
    spans.sort(key=lambda ab: -(ab[1] - ab[0]))           # stable: the chains still running are always a prefix
    starts = np.array([a for a, _b in spans], np.int32)
    lens = np.array([b - a for a, b in spans], np.int64)
    is_key_np = np.zeros(nt, np.uint8)
    is_key_np[:p] = 1
    is_key_np[starts] = 1
    pred_slot_np = np.full(nt, -1, np.int32)
    if p > 1:
        pred_slot_np[1:p] = 0
    is_key = torch.from_numpy(is_key_np).to(dev)
    pred_slot = torch.from_numpy(pred_slot_np).to(dev)
    apply_eb = torch.zeros(nt, dtype=torch.uint8, device=dev)
    key = torch.from_numpy(starts).to(dev)
    idx = key + 1
    last = torch.full_like(key, -1)
    sse = torch.zeros(len(spans), dtype=torch.float64, device=dev)
    cnt = torch.zeros(len(spans), dtype=torch.int32, device=dev)
    net.p0(out=pool[0])
    mb = net.max_batch
    X = torch.empty((len(spans), Hp, Wp, C), dtype=torch.float32, device=dev)
    cursor = 1
    for step in range(int(lens.max()) - 1):
        B = int(np.count_nonzero(lens - 1 > step))
        ops.dwp_gather(frames, pool, key, idx, last, X, B)                                 # compress.py:219 / :222
        out = pool[cursor:cursor + B]
        for b0 in range(0, B, mb):
            net.next(X[b0:min(b0 + mb, B)], out=out[b0:min(b0 + mb, B)])                   # compress.py:224-229
        sse_step = ops.window_sse(frames, idx[:B], out)                                    # compress.py:245-246
        ops.dwp_update(sse_step, key, idx, last, sse, cnt, pred_slot, apply_eb, is_key, B, cursor, denom, threshold,
                       window, p)                                                          # compress.py:249-263
        cursor += B
    return is_key, pred_slot, apply_eb, cursor


# ------------------------------------------------------------------------------------------------ compress
@dataclass
class Encoded:
    shape: tuple                 # (1, nt, H, W, C) as written in the trailer (compress.py:390-391)
    p: int
    keys: list
    key_plane: torch.Tensor      # u8 [nt,H,W,C] (device)
    body: torch.Tensor           # int16 [N] (device): ranks, or the delta stream with entropy=False
    _table: np.ndarray           # int16 [T] or None (see `table`)
    pred_slot: np.ndarray
    pool: torch.Tensor = None    # kept only when keep_pool=True
    x: torch.Tensor = None
    copies_done: object = None   # encode_frames_host(wait_copies=False): CUDA event after the device->host copies
    _pending: object = None      # defer=True: the host-side end of the entropy stage, not run yet (finalize())
    _sink: object = None

    def finalize(self):
        """defer=True records: waits for the table and the flags of THIS sequence (queued long before its last
        kernel), raises what a synchronous call would have raised, and -- only when the table needs the reference's
        chained replacement -- queues the rank map again.  Idempotent; `table` and `payload()` call it."""
        fn, self._pending = self._pending, None
        if fn is not None:
            self._table = fn()
            if self._sink is not None:
                self.copies_done = self._sink.done
        return self

    @property
    def table(self):
        self.finalize()
        return self._table

    def payload(self):
        """entropy.dat before zstd (compress.py:375-395), host int16 (int32 for 16-bit samples: container v2)."""
        return pack_payload(self.body.cpu().numpy(), self.table, self.shape, self.p)


def pack_payload_v2(body, table, shape, p, bits=16):
    """Container v2 stream (16-bit samples), int32 little-endian, the layout of compress.py:375-394 with wider fields:
    [codes N] [table T] [T | -1] [1, nt, H, W, C] [p] [bits] [version] [magic]."""
    tail = ([int(v) for v in table] + [len(table)]) if table is not None else [-1]
    tail += [int(v) for v in shape] + [int(p), int(bits), V2_VERSION, V2_MAGIC]
    if max(int(v) for v in shape) >= 2 ** 31 or int(p) >= 2 ** 31:
        raise TezipError("sequence shape %r does not fit the v2 trailer" % (tuple(shape),))
    return np.concatenate([np.asarray(body, np.int32), np.array(tail, np.int64).astype(np.int32)])


def is_v2_payload(data):
    """True if the raw bytes / array of an entropy.dat stream end with the v2 magic."""
    a = np.asarray(data)
    if a.dtype not in (np.uint8, np.int16, np.int32):
        return False
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return raw.size >= 44 and raw.size % 4 == 0 and int(raw[-4:].view("<u4")[0]) == V2_MAGIC


def pack_payload(body, table, shape, p):
    if np.asarray(body).dtype == np.int32:
        return pack_payload_v2(body, table, shape, p)
    # the trailer is int16 like the stream (compress.py:390-394): shapes beyond 32767 cannot be represented and the
    # reference would silently wrap them (decompress.py:111-113 reads them back as int16)
    if max(int(v) for v in shape) > 32767 or int(p) > 32767:
        raise TezipError("sequence shape %r (p=%d) does not fit the container's int16 trailer" % (tuple(shape), int(p)))
    tail = []
    if table is not None:
        tail += [int(v) for v in table] + [len(table)]        # compress.py:383-385
    else:
        tail += [-1]                                          # compress.py:387
    tail += [int(v) for v in shape] + [int(p)]                # compress.py:390-392
    return np.concatenate([np.asarray(body, np.int16), np.array(tail, np.int64).astype(np.int16)])  # :394


_SIDE_STREAMS = {}


def side_stream(device, kind="out"):
    """Cached copy streams per device (creating streams per call costs more than the copies it hides): one for
    device->host results ("out") and one for host->device inputs ("in"), so that the upload of the NEXT sequence never
    queues behind the download of the previous one when calls are pipelined."""
    key = (str(device), kind)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class HostSink:
    """Streams the results of an encode to pinned host buffers while the GPU keeps working: the key plane as soon as
    the schedule is known (before any PredNet step), the int16 stream in chunks as the rank-map kernel produces them.
    Copies run on a side stream; `finish()` makes the current stream wait for them.

    Key plane: nine frames in ten of it are zero (compress.py:183), so only the KEY FRAMES cross the bus when the
    schedule is known on the host (static windows): the host plane is zeroed once, and afterwards only frames that
    were keys of the previous call into the same buffer but are not any more are cleared."""

    def __init__(self, key_host, body_host, device, chunks=4, wait_copies=True):
        self.key_host, self.body_host, self.chunks = key_host, body_host, max(1, int(chunks))
        self.device = device
        self.stream = side_stream(device)
        self.wait_copies = wait_copies
        self.done = None   # wait_copies=False: event after the last copy (finish())

    def _after_current(self):
        self.stream.wait_event(torch.cuda.current_stream(self.device).record_event())

    def key_plane(self, t, keys=None):
        self._after_current()
        kh = self.key_host
        if keys is None or not kh.is_pinned() or len(keys) * 2 > t.shape[0]:
            kh._tz_key_state = None
            with torch.cuda.stream(self.stream):
                kh.view(-1).copy_(t.view(-1), non_blocking=True)
            return
        nt = t.shape[0]
        fb = t[0].numel() * t.element_size()
        ident = (kh.numel() * kh.element_size(), nt)
        old = getattr(kh, "_tz_key_state", None)     # kept on the caller's tensor object: a new buffer starts clean
        kview = kh.view(nt, -1)
        keyset = frozenset(int(k) for k in keys)
        if old is None or old[0] != ident:
            kview.zero_()                                   # first use of this buffer: one host memset
        else:
            for fidx in old[1] - keyset:
                kview[fidx].zero_()
        kh._tz_key_state = (ident, keyset)
        lib = _lib.load()
        st = ctypes.c_void_p(self.stream.cuda_stream)
        ks = sorted(keyset)
        # runs of equally spaced keys -> one strided DMA each (p, p+W, p+2W, ... is a single run)
        i = 0
        while i < len(ks):
            j = i + 1
            stride = ks[j] - ks[i] if j < len(ks) else 1
            while j + 1 < len(ks) and ks[j + 1] - ks[j] == stride:
                j += 1
            n_run = j - i + 1 if j < len(ks) else 1
            off = ks[i] * fb
            check(lib.tz_memcpy2d_async(ctypes.c_void_p(kh.data_ptr() + off), stride * fb,
                                        ctypes.c_void_p(t.data_ptr() + off), stride * fb, fb, n_run, st),
                  "tz_memcpy2d_async")
            i += n_run

    def body_chunk(self, t, a, b):
        self._after_current()
        with torch.cuda.stream(self.stream):
            self.body_host[a:b].copy_(t[a:b], non_blocking=True)

    def finish(self):
        if self.wait_copies:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        else:
            self.done = self.stream.record_event()


_FLAG_SCRATCH = {}


def _flag_scratch(device, nt):
    """Two pinned u8[nt] landing buffers per device (allocating pinned memory per call costs a cudaHostAlloc, which
    also serialises with collectives in flight)."""
    key = str(device)
    cur = _FLAG_SCRATCH.get(key)
    if cur is None or cur.shape[1] < nt:
        cur = _FLAG_SCRATCH[key] = torch.empty((2, max(nt, 1024)), dtype=torch.uint8).pin_memory()
    return cur[0, :nt], cur[1, :nt]


class _Landing:
    """Pinned landing buffers for the small results of ONE encode (table + meta, overflow count, key-frame flags).
    Cached in a ring per device -- pinning memory per call is slow -- and handed out round-robin, so that a deferred
    record (defer=True) can still read its own table while the next sequences are being queued.  A slot whose
    previous owner has not been finalised yet finalises it first."""
    RING = 4

    def __init__(self):
        self.tm = {}          # wide -> pinned table|meta buffer
        self.ovf = torch.empty(1, dtype=torch.int64).pin_memory()
        self.flags = None     # pinned u8 [2, cap]
        self.owner = None     # Encoded whose finalize() has not run yet

    def table_meta(self, wide):
        if wide not in self.tm:
            t = torch.empty(TZ_WIDE_BINS + 2, dtype=torch.int32) if wide else torch.empty(TZ_HIST_BINS + 4, dtype=torch.int16)
            self.tm[wide] = t.pin_memory()
        return self.tm[wide]

    def flag_pair(self, nt):
        if self.flags is None or self.flags.shape[1] < nt:   # generous: growing means a cudaHostAlloc (a device-wide stall)
            self.flags = torch.empty((2, max(2 * nt, 32768)), dtype=torch.uint8).pin_memory()
        return self.flags[0, :nt], self.flags[1, :nt]


_STAGING = {}


def _staging(device, name, shape, dtype, depth=3):
    """Device staging buffers of the streaming host-buffer calls: a ring of `depth` tensors per (device, purpose,
    shape, type), allocated once, with one event per slot ("busy until") that the NEXT user of the slot waits for.
    The caching allocator is deliberately not used here: a tensor that is filled on a copy stream and read on the
    compute stream can only be recycled after cross-stream events, so its blocks come free at timing-dependent
    moments and an occasional cudaMalloc lands in the middle of a stream of sequences -- seen to stall the host for
    50-170 ms.  -> (tensor, ring, slot index); the caller stores ring["busy"][slot] when it has queued its last
    reader."""
    key = (str(device), name, tuple(int(v) for v in shape), dtype)
    ring = _STAGING.get(key)
    if ring is None:
        stale = [k for k in _STAGING if k[:2] == key[:2]]       # another shape for the same purpose: drop the old ring
        if stale:
            torch.cuda.synchronize(device)                      # (copies on the side streams may still touch it)
            for k in stale:
                del _STAGING[k]
        bufs = [torch.empty(key[2], dtype=dtype, device=device) for _ in range(depth)]
        # fresh blocks may alias memory that kernels already queued on the current stream still use
        born = torch.cuda.current_stream(device).record_event()
        ring = _STAGING[key] = {"bufs": bufs, "busy": [born] * depth, "i": 0}
    i = ring["i"]
    ring["i"] = (i + 1) % depth
    return ring["bufs"][i], ring, i


_LANDING = {}


def _landing(device):
    key = str(device)
    ring = _LANDING.setdefault(key, [[_Landing() for _ in range(_Landing.RING)], 0])
    ring[1] += 1
    slot = ring[0][ring[1] % _Landing.RING]
    if slot.owner is not None:
        owner, slot.owner = slot.owner, None
        owner.finalize()
    return slot


def is_lossless(mode, bound):
    """compress.py:24,35: BOUND_VALUE[0] == 0 (or absrel with BOUND_VALUE[1] == 0) leaves diff untouched."""
    return float(bound[0]) == 0.0 or (mode == "absrel" and float(bound[1]) == 0.0)


def pool_slots_upper_bound(nt):
    return nt + 1


def encode_frames(frames, net, p, window, threshold, mode, bound, entropy=True, dwp_chains=1, keep_pool=False,
                  keep_x=False, comm=None, sink=None, frames_ready=None, arrive=None, defer=False):
    """compress.py:176-395 on a device tensor `frames` u8 [nt,H,W,C].

    comm: optional shard communicator (tezip_b200/dist.py) when `frames` is one rank's window-aligned shard of a
    longer sequence: comm.exchange_last_x(x_last) -> (has_prev, prev_x) supplies the one-element halo of the 1-D
    delta (compress.py:75 crosses shard boundaries) and comm.reduce_hist(t) sums the symbol histogram across
    ranks so that every rank derives the same table."""
    assert frames.is_cuda and frames.dtype in (torch.uint8, torch.uint16) and frames.is_contiguous() and \
        frames.dim() == 4
    nt, H, W, C = frames.shape
    dev = frames.device
    Hp, Wp = padding_size(H), padding_size(W)
    if (Hp, Wp, C) != net.frame_shape():
        raise TezipError("ERROR:Image size is out of scope for this model.")            # compress.py:178-181
    if mode not in ops.MODES:
        raise TezipError("unknown mode %r" % (mode,))
    if any(float(b) < 0 for b in bound) and mode != "abs":
        raise TezipError("error bounds must be non-negative")
    staged = None
    if threshold is None:
        plan = cached_plan(nt, p, swp_keys(nt, p, window, shard=comm is not None))
        pool = torch.empty((plan.n_slots, Hp, Wp, C), dtype=torch.float32, device=dev)
        keys, pred_slot_np, apply_np = plan.keys, plan.pred_slot, plan.apply_eb
        # the schedule is static: upload it and emit the key plane before the first PredNet step is queued, so that
        # nothing on the host waits behind the predictions and the key plane's D2H copy runs under them.  (Queuing
        # this behind the first PredNet step instead -- the GPU then never waits for these small launches -- was
        # measured: +0.4 % with device-resident frames, but -8 % through the host-buffer API, where the flag copies
        # then sit in the copy queue behind the frames still being uploaded.)
        if arrive is None:
            staged = stage_plan(frames, keys, pred_slot_np, apply_np, sink, plan)
            run_plan(net, frames, plan, pool)
        else:
            # streaming loader: `frames` fills up group by group (arrive(a, b) uploads [a, b)); the key plane needs
            # every key frame, so the schedule is staged after the last group has been queued
            run_plan(net, frames, plan, pool, arrive)
            staged = stage_plan(frames, keys, pred_slot_np, apply_np, sink, plan)
        if frames_ready is not None:   # the non-key frames were still in flight (upload_frames); the residual needs them
            torch.cuda.current_stream(dev).wait_event(frames_ready)
    else:
        if frames_ready is not None:
            torch.cuda.current_stream(dev).wait_event(frames_ready)
        pool = torch.empty((pool_slots_upper_bound(nt), Hp, Wp, C), dtype=torch.float32, device=dev)
        is_key, pred_slot, apply_dev, _n = run_dwp(net, frames, p, threshold, pool, dwp_chains, window,
                                                   shard=comm is not None)
        staged = stage_device(frames, is_key, pred_slot, apply_dev, sink)
        keys = pred_slot_np = apply_np = None          # read back once, after everything has been queued
    enc = encode_with_pool(frames, pool, pred_slot_np, apply_np, keys, p, mode, bound, entropy, keep_pool, keep_x,
                           comm, sink, staged, defer=defer and keys is not None)
    if keys is None:
        enc.keys = [int(k) for k in np.nonzero(staged[3][1].numpy())[0]]
        enc.pred_slot = staged[1].cpu().numpy()
    return enc


def stage_plan(frames, keys, pred_slot_np, apply_np, sink=None, plan=None):
    """Device copies of the schedule (pred_slot int32 [nt], apply u8 [nt]) and the key plane (compress.py:183-263).
    plan: the cached Plan these tables belong to -- its device copies are then uploaded once, not per call."""
    dev = frames.device
    nt = frames.shape[0]

    def is_key_host():
        is_key = np.zeros(nt, np.uint8)
        is_key[list(keys)] = 1
        return torch.from_numpy(is_key)

    if plan is not None:
        pred_slot = _plan_dev(plan, dev, "pred_slot", lambda: torch.from_numpy(np.ascontiguousarray(pred_slot_np, np.int32)))
        apply = _plan_dev(plan, dev, "apply", lambda: torch.from_numpy(np.ascontiguousarray(apply_np, np.uint8)))
        is_key = _plan_dev(plan, dev, "is_key", is_key_host)
        keys_sorted = plan.dev.get("keys_sorted")
        if keys_sorted is None:
            keys_sorted = plan.dev["keys_sorted"] = sorted(keys)
    else:
        pred_slot = torch.from_numpy(np.ascontiguousarray(pred_slot_np, np.int32)).to(dev)
        apply = torch.from_numpy(np.ascontiguousarray(apply_np, np.uint8)).to(dev)
        is_key = is_key_host().to(dev)
        keys_sorted = sorted(keys)
    return stage_device(frames, is_key, pred_slot, apply, sink, keys_host=keys_sorted)


def stage_device(frames, is_key, pred_slot, apply, sink=None, keys_host=None):
    """The key plane from a schedule that already lives on the device (is_key u8[nt]).  keys_host: the key list when
    the host knows it (static windows): the sink then copies only the key frames to the host."""
    dev = frames.device
    nt = frames.shape[0]
    key_plane = ops.key_plane(frames, is_key)
    if sink is not None:
        sink.key_plane(key_plane, keys_host)
    # An all-zero key frame (a fade to black) cannot be told from a non-key frame by the decoder
    # (decompress.py:123-127): the reference then silently decodes the wrong window.  The check costs one pass over
    # the key plane and two tiny asynchronous copies; encode_with_pool looks at the answer once everything is queued.
    nz = ops.frames_nonzero(key_plane)
    landing = _landing(dev)                       # cached pinned buffers: consumed by check_key_frames() / finalize()
    nz_host, ik_host = landing.flag_pair(nt)
    # The two flag copies run on the device->host side stream, never on the compute stream: in streaming use
    # (encode_frames_host(wait_copies=False)) the previous sequence's 123 MB stream may still be draining through the
    # same copy engine, and a copy queued on the compute stream would hold back every PredNet kernel behind it.
    main, out = torch.cuda.current_stream(dev), side_stream(dev)
    out.wait_event(main.record_event())
    with torch.cuda.stream(out):
        nz_host.copy_(nz, non_blocking=True)
        ik_host.copy_(is_key, non_blocking=True)
        done = out.record_event()
    nz.record_stream(out)
    is_key.record_stream(out)
    return pred_slot, apply, key_plane, (nz_host, ik_host, done, landing)


def check_key_frames(staged):
    """Raises if a scheduled key frame is all zero (see stage_plan)."""
    nz_host, ik_host, ev = staged[3][:3]
    ev.synchronize()
    bad = np.nonzero((nz_host.numpy() == 0) & (ik_host.numpy() != 0))[0]
    if bad.size:
        raise TezipError("frame %d is a key frame but all-zero: the container cannot mark it (decompress.py:123-127 "
                         "finds key frames as 'not all-zero'); drop or perturb the frame" % int(bad[0]))


def encode_with_pool(frames, pool, pred_slot_np, apply_np, keys, p, mode, bound, entropy=True, keep_pool=False,
                     keep_x=False, comm=None, sink=None, staged=None, defer=False):
    """compress.py:271-395 given the predictions: key plane, residual, error bound, delta, table, rank map.
    staged: the result of stage_plan() when the caller already ran it (before the predictions).
    defer: return as soon as everything is queued; the host-side end of the entropy stage (reading the table, the
    overflow / collision / key-frame flags) runs in Encoded.finalize() -- a streaming caller queues the next sequence
    first, so the GPU never waits for the host between sequences.
    frames u8 -> the reference's int16 stream; frames u16 -> the container-v2 int32 stream (same steps, wider codes)."""
    nt, H, W, C = frames.shape
    dev = frames.device
    wide = ops.is_wide(frames)
    nbins = TZ_WIDE_BINS if wide else TZ_HIST_BINS
    code_dtype = torch.int32 if wide else torch.int16
    if staged is None:
        staged = stage_plan(frames, keys, pred_slot_np, apply_np, sink)
    pred_slot, apply_dev, key_plane = staged[:3]
    N = nt * H * W * C
    body = torch.empty(N, dtype=code_dtype, device=dev)
    table = None
    x = None
    lossless = is_lossless(mode, bound)
    # lossy, 8-bit, plane-wide bound: ONE data pass computes residual + error bound + delta histogram (tz_encode_lossy)
    fused_lossy = (not lossless) and entropy and ops.encode_lossy_supported(frames, mode)
    hist_ovf = torch.zeros(nbins + 2, dtype=torch.int64, device=dev) if entropy else None   # hist | overflow | counter
    if fused_lossy:
        x = ops.encode_lossy(frames, pool, pred_slot, apply_dev, mode, list(bound), hist_ovf[:nbins],
                             hist_ovf[nbins:nbins + 1], hist_ovf[nbins + 1:].view(torch.int32),
                             has_prev=3 if comm is not None else 0)
    elif not lossless or keep_x:
        x = ops.residual(frames, pool, pred_slot)                                        # compress.py:293-314
        if not lossless:
            ops.error_bound(frames, x, apply_dev, mode, list(bound))                    # compress.py:315-319
    has_prev, prev_x = False, None
    if comm is not None:
        # the one-element halo of the 1-D delta (compress.py:75 crosses shard boundaries) never visits the host: this
        # shard's last residual is left in a device int32, all-gathered, and the kernels read element rank-1
        if x is not None:
            x_last = x.view(-1)[-1:].to(torch.int32)
        else:   # lossless: the last residual straight from the last frame
            x_last = ops.last_residual(frames, pool, pred_slot)
        has_prev, prev_x = comm.exchange_last_x(x_last)

    def hist_pass(hist, ovf):
        if fused_lossy:
            if comm is not None:     # the stream's first symbol, now that the halo is here
                ops.finding_difference_hist(x.view(-1)[:1], hist, ovf, has_prev, prev_x)
        elif wide:
            ops.encode16(frames, pool, pred_slot, x, 0, hist=hist, overflow=ovf, has_prev=has_prev, prev_x=prev_x)
        elif x is not None:
            ops.finding_difference_hist(x, hist, ovf, has_prev, prev_x)                  # :339-340,348-355
        else:
            ops.encode_lossless(frames, pool, pred_slot, 0, hist=hist, overflow=ovf, has_prev=has_prev,
                                prev_x=prev_x)

    if entropy:
        hist, ovf = hist_ovf[:nbins], hist_ovf[nbins:nbins + 1]            # one buffer: one D2H, one reduce
        hist_pass(hist, ovf)
        if comm is not None:
            comm.reduce_hist(hist_ovf[:nbins + 1])
        # table and symbol -> rank LUT on the device: the rank-map pass is queued right behind the histogram pass,
        # the host reads the table (and the overflow / collision flags) only after everything has been launched
        if wide:
            tm = torch.empty(nbins + 2, dtype=torch.int32, device=dev)      # table | meta (2 x int32)
            table_dev, meta = tm[:nbins], tm[nbins:]
            lut = torch.empty(nbins, dtype=torch.int32, device=dev)
            ops.build_table16_device(hist, table_dev, lut, meta)
        else:
            tm = torch.empty(nbins + 4, dtype=torch.int16, device=dev)      # table | meta (2 x int32)
            table_dev, meta = tm[:nbins], tm[nbins:].view(torch.int32)
            lut = torch.empty(nbins, dtype=torch.int16, device=dev)
            ops.build_table_device(hist, table_dev, lut, meta)                            # :352-361, :84-90
        # their (small) copies to the host are queued NOW, ahead of the stream's large device->host copies in the
        # copy engine's queue; they are waited for at the end
        landing = staged[3][3]
        tm_host, ovf_host = landing.table_meta(wide), landing.ovf
        tm_host.copy_(tm, non_blocking=True)
        ovf_host.copy_(ovf, non_blocking=True)
        small_ready = torch.cuda.current_stream(dev).record_event()
    else:
        lut = None

    def rank_call(a, b, lut_, hp, px):
        """codes of stream elements [a, b) (a, b multiples of 8 or the ends)."""
        if wide:
            if x is not None:
                ops.finding_difference_rank16(x.view(-1)[a:b], lut_, body[a:b], has_prev=hp, prev_x=px)
            else:
                assert a == 0 and b == N
                ops.encode16(frames, pool, pred_slot, None, 1, lut=lut_, out=body, has_prev=hp, prev_x=px)
        elif x is not None:
            ops.finding_difference_rank(x.view(-1)[a:b], lut_, out=body[a:b], has_prev=hp, prev_x=px)  # :339-340,369
        else:
            assert a == 0 and b == N
            ops.encode_lossless(frames, pool, pred_slot, 1, lut=lut_, out=body, has_prev=hp, prev_x=px)

    def rank_pass(lut_):
        if x is not None and sink is not None and sink.chunks > 1:
            # rank map in chunks so that the device->host copy of chunk i overlaps the kernel of chunk i+1
            step = -(-N // sink.chunks) // 8 * 8 + 8
            for a in range(0, N, step):
                b = min(N, a + step)
                if a == 0:
                    rank_call(a, b, lut_, has_prev, prev_x)
                else:
                    rank_call(a, b, lut_, 2, None)                 # y[a] = x[a-1] - x[a]
                sink.body_chunk(body, a, b)
        else:
            rank_call(0, N, lut_, has_prev, prev_x)
            if sink is not None:
                sink.body_chunk(body, 0, N)

    rank_pass(lut)

    def host_end():
        """The host-side end of the entropy stage: -> table.  (The small copies it waits for were queued ahead of the
        stream's large device->host copies.)"""
        table = None
        staged[3][3].owner = None
        if entropy:
            small_ready.synchronize()
            if int(ovf_host[0]) != 0:
                raise TezipError("residual symbols fall outside [0, %d): the reference's bincount/int16 stream "
                                 "cannot represent this bound" % nbins)
            if wide:
                table = tm_host[:int(tm_host[nbins])].numpy().copy()
            else:
                meta_np = tm_host[TZ_HIST_BINS:].view(torch.int32).numpy()
                table = tm_host[:int(meta_np[0])].numpy().copy()
                if int(meta_np[1]) != 0:   # a symbol inside the rank range: the reference's sequential replacement chains
                    rank_pass(torch.from_numpy(ops.encode_lut(table)).to(dev))
                    if sink is not None:
                        sink.finish()
                    if defer:   # (rare) the repeated pass may re-read a staging buffer that is about to be recycled
                        torch.cuda.current_stream(dev).synchronize()
        check_key_frames(staged)
        return table

    enc = Encoded((1, nt, H, W, C), p, None if keys is None else list(keys), key_plane, body, None,
                  None if pred_slot_np is None else np.asarray(pred_slot_np), pool if keep_pool else None,
                  x if keep_x else None)
    enc._sink = sink
    if defer:
        if sink is not None:
            sink.finish()
        enc._pending = host_end
        staged[3][3].owner = enc
    else:
        enc._table = host_end()
        if sink is not None:
            sink.finish()
    return enc


# ------------------------------------------------------------------------------------------------ decompress
def parse_payload_v2(data):
    data = np.asarray(data).view(np.uint8).reshape(-1).view("<i4")
    if data.size < 11 or (int(data[-1]) & 0xffffffff) != V2_MAGIC or int(data[-2]) != V2_VERSION:
        raise TezipError("not a container-v2 stream")
    bits, p = int(data[-3]), int(data[-4])
    if bits != 16:
        raise TezipError("container v2 with %d-bit samples is not supported" % bits)
    shape = tuple(int(v) for v in data[-9:-4])
    data = data[:-9]
    table_len = int(data[-1])
    if table_len == -1:
        return data[:-1], None, shape, p
    if table_len < 0 or table_len + 1 > data.size:
        raise TezipError("corrupt entropy.dat trailer (table length %d)" % table_len)
    table_start = data.size - table_len - 1
    return data[:table_start], data[table_start:-1].copy(), shape, p


def parse_payload(data):
    """decompress.py:103-113,203-221 -> (body view int16, table or None, shape(5), p); container-v2 streams
    (recognised by their magic) -> the same with int32 body / table."""
    if is_v2_payload(data):
        return parse_payload_v2(data)
    data = np.asarray(data, dtype=np.int16)
    if data.size < 8:
        raise TezipError("entropy.dat payload is too short")
    p = int(data[-1])
    shape = tuple(int(v) for v in data[-6:-1])
    data = data[:-6]
    table_len = int(data[-1])
    if table_len == -1:
        return data[:-1], None, shape, p
    if table_len < 0 or table_len + 1 > data.size:
        raise TezipError("corrupt entropy.dat trailer (table length %d)" % table_len)
    table_start = data.size - table_len - 1
    return data[:table_start], data[table_start:-1].copy(), shape, p


def decode_arrays(key_plane, body, table, shape, p, net, want_x=False, first_mode=0, first_x=0, body_event=None,
                  nonzero=None, out=None):
    """decompress.py:115-256,269 on device tensors: key_plane u8 [nt,H,W,C], body int16 [N] -> u8 frames.
    body_event: optional CUDA event after which `body` is valid (its host->device copy may still be in flight on
    another stream while the predictions are replayed; only the final reconstruct needs it).
    nonzero: optional host u8[nt], the key-frame flags of decompress.py:123-127 when the caller has already computed
    them (decode_arrays_host does, on its upload stream).  out: optional preallocated result tensor."""
    _one, nt, H, W, C = shape
    dev = key_plane.device
    Hp, Wp = padding_size(H), padding_size(W)
    if (Hp, Wp, C) != net.frame_shape():
        raise TezipError("ERROR:keyframe size and model size do not match.")             # decompress.py:131-135
    if body.numel() != nt * H * W * C:
        raise TezipError("entropy.dat holds %d residuals, shape needs %d" % (body.numel(), nt * H * W * C))
    key_plane = key_plane.view(nt, H, W, C)
    nz = ops.frames_nonzero(key_plane).cpu().numpy() if nonzero is None else np.asarray(nonzero)   # decompress.py:123-127
    keys = [int(i) for i in np.nonzero(nz)[0]]
    plan = cached_plan(nt, p, keys)
    pool = torch.empty((plan.n_slots, Hp, Wp, C), dtype=torch.float32, device=dev)
    pred_slot = _plan_dev(plan, dev, "pred_slot", lambda: torch.from_numpy(np.ascontiguousarray(plan.pred_slot, np.int32)))
    if ops.is_wide(body) != ops.is_wide(key_plane):
        raise TezipError("key plane and stream disagree about the sample width")
    if table is not None:
        lut = torch.from_numpy(ops.decode_lut16(table) if ops.is_wide(body) else ops.decode_lut(table)).to(dev)
        tl = len(table)
    else:
        lut, tl = None, -1
    run_plan(net, key_plane, plan, pool)                                                 # decompress.py:138-189
    if body_event is not None:
        torch.cuda.current_stream(dev).wait_event(body_event)
    return ops.reconstruct(body, (nt, H, W, C), Hp, Wp, tl, lut, pool, pred_slot, key_plane, first_mode, first_x,
                           want_x=want_x, out=out), plan


def encode_frames_host(frames_host, net, p, window, threshold, mode, bound, key_host, body_host, entropy=True,
                       dwp_chains=1, comm=None, chunks=4, wait_copies=True, defer=False):
    """Host-buffer API: frames_host u8 [nt,H,W,C] (pinned) -> key_host u8 (pinned, same shape), body_host int16 [N]
    (pinned).  The H2D copy, the kernels and the D2H copies are pipelined; returns the Encoded record (table, keys)
    after the copies have been ordered on the current stream (synchronise before reading the host buffers).
    wait_copies=False (streaming use: the next sequence's kernels should not queue behind this one's device->host
    copies): the copies are only ordered on the side stream and the record carries `copies_done`, the event to
    synchronise before reading key_host / body_host.  Keep the returned record alive until then (it owns the device
    tensors the copies read) and give consecutive calls different host buffers.
    defer=True (static windows): return without waiting for this sequence's table; call `finalize()` on the record
    (or read its `table`) after the NEXT sequence has been queued -- the GPU then never idles between sequences."""
    dev = net.device
    sink = HostSink(key_host, body_host, dev, chunks, wait_copies)
    if defer and threshold is None and frames_host.is_pinned():
        # streaming: the host runs a sequence ahead of the GPU, so the whole upload goes to the upload stream NOW (it
        # lands while the previous sequence is still being predicted) instead of key-frames-first on the compute stream
        main, side = torch.cuda.current_stream(dev), side_stream(dev, "in")
        frames, ring, slot = _staging(dev, "frames_in", frames_host.shape, frames_host.dtype)
        side.wait_event(ring["busy"][slot])    # the sequence that last used this staging buffer has been encoded
        with torch.cuda.stream(side):
            frames.copy_(frames_host, non_blocking=True)
            landed = side.record_event()
        main.wait_event(landed)
        ready = None
    else:
        frames, ready, ring = upload_frames(frames_host, dev, p, window, threshold) + (None,)
    enc = encode_frames(frames, net, p, window, threshold, mode, bound, entropy, dwp_chains, comm=comm, sink=sink,
                        frames_ready=ready, defer=defer)
    if ring is not None:
        ring["busy"][slot] = torch.cuda.current_stream(dev).record_event()   # every reader of `frames` is queued
    enc.copies_done = sink.done
    return enc


def upload_frames(frames_host, dev, p, window, threshold):
    """Host (pinned) -> device copy of the frames.  With a static window (SWP, p = 0) the key frames go first, on
    the current stream: the prediction steps read nothing else (compress.py:219-229).  The other frames follow on
    the side stream, behind the PredNet kernels; the returned event marks their arrival (None: everything was
    copied on the current stream)."""
    nt = frames_host.shape[0]
    if threshold is not None or p != 0 or window is None or window < 2 or nt <= window or not frames_host.is_pinned():
        return frames_host.to(dev, non_blocking=True), None
    lib = _lib.load()
    fb = frames_host[0].numel() * frames_host.element_size()
    n_full, rem = divmod(nt, window)
    main = torch.cuda.current_stream(dev)
    frames = torch.empty(frames_host.shape, dtype=frames_host.dtype, device=dev)
    src, dst, pitch = frames_host.data_ptr(), frames.data_ptr(), window * fb

    def copy2d(offset, width, height, stream):
        check(lib.tz_memcpy2d_async(ctypes.c_void_p(dst + offset), pitch, ctypes.c_void_p(src + offset), pitch, width,
                                    height, ctypes.c_void_p(stream.cuda_stream)), "tz_memcpy2d_async")

    copy2d(0, fb, n_full + (1 if rem else 0), main)                 # first frame of every window
    side = side_stream(dev, "in")
    side.wait_event(main.record_event())                            # `frames` may reuse memory the main stream still reads
    copy2d(fb, (window - 1) * fb, n_full, side)                     # frames 1..window-1 of the full windows
    if rem > 1:
        copy2d(n_full * pitch + fb, (rem - 1) * fb, 1, side)        # and of the trailing short one
    return frames, side.record_event()


def decode_arrays_host(key_host, body_host, table, shape, p, net, out_host, first_mode=0, first_x=0,
                       wait_copies=True):
    """Host-buffer API of the decoder: the key plane goes first (the prediction replay needs it), the int16 stream
    follows while PredNet runs, the frames come back at the end.  Both uploads and the key-frame scan
    (decompress.py:123-127) run on the upload stream and the download on the download stream, so in streaming use
    (wait_copies=False, consecutive sequences, alternating out_host buffers) the next sequence's key plane is already
    on the device -- and its schedule known to the host -- while the current sequence is still being predicted, and
    its kernels never queue behind the current sequence's download.  wait_copies=False returns (out, plan, event):
    synchronise the event before reading out_host."""
    dev = net.device
    main = torch.cuda.current_stream(dev)
    _one, nt, H, W, C = shape
    s_in = side_stream(dev, "in")
    # device staging from rings (see _staging): key plane and stream are filled on the upload stream, read on the
    # compute stream; the result is written on the compute stream and read by the download stream
    key_plane, kring, kslot = _staging(dev, "key_in", key_host.shape, key_host.dtype)
    body, bring, bslot = _staging(dev, "body_in", body_host.shape, body_host.dtype)
    s_in.wait_event(kring["busy"][kslot])
    s_in.wait_event(bring["busy"][bslot])
    with torch.cuda.stream(s_in):
        key_plane.copy_(key_host, non_blocking=True)
        nz_dev = ops.frames_nonzero(key_plane.view(nt, H, W, C))
        nz_host = _flag_scratch(dev, nt)[0]
        nz_host.copy_(nz_dev, non_blocking=True)
        ev_key = s_in.record_event()
        body.copy_(body_host, non_blocking=True)
        ev_body = s_in.record_event()
    ev_key.synchronize()               # the host needs the key positions to build the schedule
    nz = nz_host.numpy().copy()
    main.wait_event(ev_key)
    out_dev = oring = None
    if not wait_copies:
        out_dev, oring, oslot = _staging(dev, "frames_out", (nt, H, W, C), key_host.dtype)
        main.wait_event(oring["busy"][oslot])      # its previous contents have been downloaded
    out, plan = decode_arrays(key_plane.view(nt, H, W, C), body.view(-1), table, shape, p, net, first_mode=first_mode,
                              first_x=first_x, body_event=ev_body, nonzero=nz, out=out_dev)
    kring["busy"][kslot] = bring["busy"][bslot] = main.record_event()   # every reader of the two inputs is queued
    if wait_copies:
        out_host.copy_(out, non_blocking=True)
        return out, plan
    s_out = side_stream(dev, "out")
    s_out.wait_event(main.record_event())
    with torch.cuda.stream(s_out):
        out_host.copy_(out, non_blocking=True)
        done = s_out.record_event()
    oring["busy"][oslot] = done
    return out, plan, done
