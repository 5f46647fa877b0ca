"""The container's lossless back-end on the GPU: compress.py:276,398 call
`zstd.compress(bytes, 9)` on the key plane and on the packed stream, decompress.py:89,98 call `zstd.decompress`.
All the reference's decoder needs is ONE valid zstd frame that carries its content size, so the frame is written
by CUDA kernels (csrc/tz_zstd.cu) from data that is already in HBM, in the subset of the format (RFC 8878) whose
blocks do not depend on each other: per 128 KB block an RLE block (the zero frames of the key plane cost 4 bytes),
a raw block, or a compressed block = the block's bytes Huffman-coded as four literal streams + zero sequences.  No
match finding: the ratio is that of an order-0 byte coder, the rate that of a memory pass (bench.py `container`
quotes both beside libzstd level 9, which stays the default: TEZIP_ZSTD_LEVEL=gpu selects this writer).

Host side (this file): the Huffman code of the frame from the device's byte histogram (length-limited to 11 bits,
canonical order of the zstd decoder), its tree description (weights, FSE-compressed as HUF_compressWeights does or
4-bit direct), and the calls.  The kernels do the rest; nothing here touches the data."""
import heapq
import struct

import numpy as np

BLOCK = 131072
MAX_BITS = 11                 # Max_Number_of_Bits of a literals Huffman code (RFC 8878 4.2.1)
FRAME_HEADER = 14
_FSE_LOG = 6                  # accuracy log of the weights' FSE table (the format's maximum for weights)


def frame_header(n):
    """magic, descriptor 0xC0 (8-byte content size), window descriptor 0x38 (128 KB), content size."""
    return b"\x28\xb5\x2f\xfd\xc0\x38" + int(n).to_bytes(8, "little")


def empty_frame():
    return frame_header(0) + b"\x01\x00\x00"          # one empty raw block, last


def _huffman_lengths(counts):
    """Plain Huffman code lengths of `counts` (all > 0, at least two)."""
    n = len(counts)
    heap = [(int(c), i) for i, c in enumerate(counts)]
    heapq.heapify(heap)
    parent = [-1] * (2 * n - 1)
    nxt = n
    while len(heap) > 1:
        c1, a = heapq.heappop(heap)
        c2, b = heapq.heappop(heap)
        parent[a] = parent[b] = nxt
        heapq.heappush(heap, (c1 + c2, nxt))
        nxt += 1
    depth = [0] * (2 * n - 1)
    for i in range(2 * n - 3, -1, -1):
        depth[i] = depth[parent[i]] + 1
    return np.array(depth[:n], np.int64)


def _limit_lengths(lens, counts, max_bits):
    """Code lengths clamped to max_bits and repaired to a complete code (Kraft sum exactly 1): symbols are lengthened
    cheapest first (smallest count) until the clamped code fits, then the slack that the last step left is given back
    to the most frequent symbols that can take it.  None if the repair does not close (the caller flattens instead)."""
    lens = [min(int(v), max_bits) for v in lens]
    full = 1 << max_bits
    excess = sum(1 << (max_bits - v) for v in lens) - full
    by_count = sorted(range(len(lens)), key=lambda i: (int(counts[i]), i))
    at = 0
    while excess > 0:
        while at < len(by_count) and lens[by_count[at]] >= max_bits:
            at += 1
        if at == len(by_count):
            return None
        i = by_count[at]
        excess -= 1 << (max_bits - lens[i] - 1)
        lens[i] += 1
    slack = -excess
    for i in reversed(by_count):                       # most frequent first
        while slack and lens[i] > 1 and (1 << (max_bits - lens[i])) <= slack:
            slack -= 1 << (max_bits - lens[i])
            lens[i] -= 1
        if not slack:
            break
    return np.array(lens, np.int64) if slack == 0 else None


def code_lengths(hist, max_bits=MAX_BITS):
    """hist: 256 counts -> 256 code lengths (0 = absent), complete prefix code with lengths <= max_bits; None if
    fewer than two byte values occur (such data is RLE, not Huffman)."""
    hist = np.asarray(hist, np.int64)
    sym = np.nonzero(hist)[0]
    if len(sym) < 2:
        return None
    counts = hist[sym].copy()
    lens = _huffman_lengths(counts)
    if lens.max() > max_bits:
        fixed = _limit_lengths(lens, counts, max_bits)
        while fixed is None:                              # (not seen in practice) flatten until the plain code fits
            counts = (counts + 1) // 2
            lens = _huffman_lengths(counts)
            fixed = lens if lens.max() <= max_bits else None
        lens = fixed
    out = np.zeros(256, np.int64)
    out[sym] = lens
    return out


def weights_and_codes(lens):
    """-> (weights[256], ctable u32[256] = code | nbits << 16).  Weight = tableLog + 1 - length; codes in the order in
    which zstd's decoder fills its table: weights ascending (longest codes first, from code 0), byte values ascending."""
    lens = np.asarray(lens, np.int64)
    table_log = int(lens.max())
    w = np.where(lens > 0, table_log + 1 - lens, 0)
    ct = np.zeros(256, np.uint32)
    pos = 0
    for wt in range(1, table_log + 1):
        for s in np.nonzero(w == wt)[0]:
            ct[s] = (pos >> (wt - 1)) | (int(lens[s]) << 16)
            pos += 1 << (wt - 1)
    assert pos == 1 << table_log, "code is not complete"
    return w, ct


# ---- FSE coding of the weights (zstd's HUF_compressWeights: FSE_normalizeCount / FSE_writeNCount / FSE_buildCTable /
# FSE_compress_usingCTable restated from the published format, RFC 8878 4.1 and 4.2.1.2) ---------------------------------

def _normalize(count, total, log):
    """Normalised counts with sum 2^log and >= 1 for every symbol that occurs."""
    size = 1 << log
    norm = [0] * len(count)
    for s, c in enumerate(count):
        if c:
            norm[s] = max(1, int(round(c * size / float(total))))
    while sum(norm) != size:
        diff = size - sum(norm)
        if diff > 0:
            norm[int(np.argmax(norm))] += diff
        else:
            s = int(np.argmax(norm))
            step = min(-diff, norm[s] - 1)
            if step == 0:
                return None
            norm[s] -= step
    return norm


def _write_ncount(norm, log):
    size = 1 << log
    bits, nbits = log - 5, 4
    remaining, threshold, nb = size + 1, size, log + 1
    symbol, alphabet, previous0 = 0, len(norm), False
    while symbol < alphabet and remaining > 1:
        if previous0:
            start = symbol
            while symbol < alphabet and norm[symbol] == 0:
                symbol += 1
            if symbol == alphabet:
                break
            while symbol >= start + 24:
                start += 24
                bits |= 0xFFFF << nbits
                nbits += 16
            while symbol >= start + 3:
                start += 3
                bits |= 3 << nbits
                nbits += 2
            bits |= (symbol - start) << nbits
            nbits += 2
        count = norm[symbol]
        symbol += 1
        mx = (2 * threshold - 1) - remaining
        remaining -= abs(count)
        count += 1
        if count >= threshold:
            count += mx
        bits |= count << nbits
        nbits += nb
        nbits -= 1 if count < mx else 0
        previous0 = count == 1
        if remaining < 1:
            return None
        while remaining < threshold:
            nb -= 1
            threshold >>= 1
    if remaining != 1:
        return None
    return bits.to_bytes((nbits + 7) // 8, "little")


def _build_ctable(norm, log):
    size = 1 << log
    mask, step = size - 1, (size >> 1) + (size >> 3) + 3
    cumul = [0]
    for c in norm:
        cumul.append(cumul[-1] + c)
    spread = [0] * size
    pos = 0
    for s, c in enumerate(norm):
        for _ in range(c):
            spread[pos] = s
            pos = (pos + step) & mask
    assert pos == 0
    state_table = [0] * size
    cur = list(cumul)
    for u in range(size):
        s = spread[u]
        state_table[cur[s]] = size + u
        cur[s] += 1
    delta_bits, delta_state = [0] * len(norm), [0] * len(norm)
    total = 0
    for s, c in enumerate(norm):
        if c == 0:
            delta_bits[s] = ((log + 1) << 16) - size
        elif c == 1:
            delta_bits[s] = (log << 16) - size
            delta_state[s] = total - 1
            total += 1
        else:
            max_bits_out = log - ((c - 1).bit_length() - 1)
            delta_bits[s] = (max_bits_out << 16) - (c << max_bits_out)
            delta_state[s] = total - c
            total += c
    return state_table, delta_bits, delta_state


def fse_compress_weights(w):
    """w: the weights of byte values 0 .. last-1 -> FSE table description + bit stream, or None when FSE cannot code
    them (fewer than two weights, or all equal)."""
    w = [int(v) for v in w]
    n = len(w)
    if n < 2:
        return None
    count = [0] * (max(w) + 1)
    for v in w:
        count[v] += 1
    if max(count) == n:
        return None
    log = _FSE_LOG
    norm = _normalize(count, n, log)
    if norm is None:
        return None
    head = _write_ncount(norm, log)
    if head is None:
        return None
    state_table, delta_bits, delta_state = _build_ctable(norm, log)
    acc = [0, 0]                                       # bit container (LSB first), bit count

    def add(value, nb):
        acc[0] |= (value & ((1 << nb) - 1)) << acc[1]
        acc[1] += nb

    def init(sym):                                     # FSE_initCState2
        nb_out = (delta_bits[sym] + (1 << 15)) >> 16
        value = (nb_out << 16) - delta_bits[sym]
        return state_table[(value >> nb_out) + delta_state[sym]]

    def encode(state, sym):                            # FSE_encodeSymbol
        nb_out = (state + delta_bits[sym]) >> 16
        add(state, nb_out)
        return state_table[(state >> nb_out) + delta_state[sym]]

    ip = n
    if n & 1:
        s1 = init(w[ip - 1]); s2 = init(w[ip - 2]); s1 = encode(s1, w[ip - 3])
        ip -= 3
    else:
        s2 = init(w[ip - 1]); s1 = init(w[ip - 2])
        ip -= 2
    while ip > 0:
        s2 = encode(s2, w[ip - 1])
        s1 = encode(s1, w[ip - 2])
        ip -= 2
    add(s2, log)
    add(s1, log)
    add(1, 1)                                          # end mark
    return head + acc[0].to_bytes((acc[1] + 7) // 8, "little")


def tree_description(weights):
    """Huffman_Tree_Description (RFC 8878 4.2.1) of weights[256]: the weights of every byte value below the largest one
    that occurs (its own weight is implied).  FSE-compressed when that is possible and shorter, else 4 bits per weight
    (at most 128 of them), else None: no tree can be written and the blocks stay raw."""
    weights = np.asarray(weights)
    last = int(np.nonzero(weights)[0].max())
    w = [int(v) for v in weights[:last]]
    direct = None
    if 1 <= len(w) <= 128:
        padded = w + [0] * (len(w) & 1)
        direct = bytes([127 + len(w)]) + bytes((padded[i] << 4) | padded[i + 1] for i in range(0, len(padded), 2))
    fse = fse_compress_weights(w)
    if fse is not None and 1 < len(fse) < 128 and (direct is None or len(fse) + 1 < len(direct)):
        return bytes([len(fse)]) + fse
    return direct


def huffman_tables(hist):
    """-> (ctable u32[256], tree bytes); (zeros, b"") when the bytes cannot be Huffman-coded."""
    lens = code_lengths(hist)
    if lens is not None:
        w, ct = weights_and_codes(lens)
        tree = tree_description(w)
        if tree is not None:
            return ct, tree
    return np.zeros(256, np.uint32), b""


# ---- reading frames back (decompress.py:89,98): the host walks the headers, the device decodes the blocks ---------------

DBLOCK = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("type", "<u4"), ("regen", "<u4"), ("stream_bytes", "<u4", 4),
                   ("table", "<u4"), ("pad", "<u4")])          # struct ZsDBlock of csrc/tz_zstd_core.h
DLOG = 11


def _read_ncount(data):
    """FSE table description (RFC 8878 4.1.1) -> (normalised counts, accuracy log, bytes used)."""
    bits = int.from_bytes(data, "little")
    log = (bits & 0xF) + 5
    bits >>= 4
    used = 4
    if log > 6:
        raise ValueError("accuracy log of the weights exceeds 6")
    remaining, threshold, nb = (1 << log) + 1, 1 << log, log + 1
    norm, previous0 = [], False
    while remaining > 1 and len(norm) < 256:
        if previous0:
            n0 = 0
            while bits & 0xFFFF == 0xFFFF:
                n0 += 24
                bits >>= 16
                used += 16
            while bits & 3 == 3:
                n0 += 3
                bits >>= 2
                used += 2
            n0 += bits & 3
            bits >>= 2
            used += 2
            norm += [0] * n0
        mx = (2 * threshold - 1) - remaining
        if (bits & (threshold - 1)) < mx:
            count = bits & (threshold - 1)
            take = nb - 1
        else:
            count = bits & (2 * threshold - 1)
            if count >= threshold:
                count -= mx
            take = nb
        bits >>= take
        used += take
        count -= 1
        remaining -= abs(count)
        norm.append(count)
        previous0 = count == 0
        while remaining < threshold:
            nb -= 1
            threshold >>= 1
    if remaining != 1 or (used + 7) // 8 > len(data):
        raise ValueError("corrupt FSE table description")
    return norm, log, (used + 7) // 8


def fse_decompress_weights(data):
    """Inverse of fse_compress_weights (zstd's FSE_decompress on Huffman weights, RFC 8878 4.2.1.2)."""
    norm, log, used = _read_ncount(data)
    size = 1 << log
    high = size - 1
    symbol = [0] * size
    nxt = []
    for s, c in enumerate(norm):
        if c == -1:
            symbol[high] = s
            high -= 1
            nxt.append(1)
        else:
            nxt.append(c)
    step, pos = (size >> 1) + (size >> 3) + 3, 0
    for s, c in enumerate(norm):
        for _ in range(max(c, 0)):
            symbol[pos] = s
            pos = (pos + step) & (size - 1)
            while pos > high:
                pos = (pos + step) & (size - 1)
    nbits, base = [0] * size, [0] * size
    for u in range(size):
        st = nxt[symbol[u]]
        nxt[symbol[u]] += 1
        nbits[u] = log - (st.bit_length() - 1)
        base[u] = (st << nbits[u]) - size
    stream = data[used:]
    if not stream or stream[-1] == 0:
        raise ValueError("corrupt FSE stream")
    value = int.from_bytes(stream, "little")
    left = (len(stream) - 1) * 8 + stream[-1].bit_length() - 1      # bits below the end mark

    def read(n):                # the n bits below the cursor; reading past the start gives zeros and a negative `left`
        nonlocal left
        left -= n
        return (value >> left) & ((1 << n) - 1) if left >= 0 else (value << -left) & ((1 << n) - 1)

    s1 = read(log)
    s2 = read(log)
    out = []
    while len(out) < 255:
        out.append(symbol[s1])
        s1 = base[s1] + read(nbits[s1])
        if left < 0:
            out.append(symbol[s2])
            break
        out.append(symbol[s2])
        s2 = base[s2] + read(nbits[s2])
        if left < 0:
            out.append(symbol[s1])
            break
    else:
        raise ValueError("corrupt FSE stream")
    return out


def decode_table(tree):
    """Huffman_Tree_Description bytes -> (decoding table u16[2^11]: symbol | nbits << 8, bytes of the description)."""
    head = tree[0]
    if head >= 128:
        count = head - 127
        used = 1 + (count + 1) // 2
        raw = tree[1:used]
        if len(raw) < used - 1:
            raise ValueError("truncated tree description")
        w = []
        for b in raw:
            w += [b >> 4, b & 15]
        w = w[:count]
    else:
        used = 1 + head
        if head == 0 or len(tree) < used:
            raise ValueError("truncated tree description")
        w = fse_decompress_weights(bytes(tree[1:used]))
    total = sum((1 << (v - 1)) for v in w if v)
    if total == 0 or len(w) > 255:
        raise ValueError("corrupt Huffman weights")
    log = total.bit_length()                      # the implied last weight completes the sum to the next power of two
    rest = (1 << log) - total
    if log > DLOG or rest & (rest - 1) or rest == 0:
        raise ValueError("corrupt Huffman weights")
    w = w + [rest.bit_length()]
    table = np.zeros(1 << DLOG, np.uint16)
    pos = 0
    for wt in range(1, log + 1):                  # the order of weights_and_codes
        span = (1 << (wt - 1)) << (DLOG - log)
        for s, v in enumerate(w):
            if v == wt:
                table[pos:pos + span] = s | ((log + 1 - wt) << 8)
                pos += span
    if pos != 1 << DLOG:
        raise ValueError("corrupt Huffman weights")
    return table, used


def parse_frame(data):
    """data: bytes / u8 array holding ONE zstd frame.  -> (content size, blocks as DBLOCK array, tables u16[T, 2048]) when
    every block is raw, RLE, or 4-stream Huffman literals with zero sequences (what compress_device writes); None for
    any other frame (the caller decodes those with libzstd, like the reference)."""
    d = memoryview(data).cast("B") if not isinstance(data, np.ndarray) else \
        memoryview(np.ascontiguousarray(data).view(np.uint8).reshape(-1))
    n = len(d)
    if n < 9 or bytes(d[:4]) != b"\x28\xb5\x2f\xfd":
        return None
    fhd = int(d[4])
    fcs_flag, single, checksum, dict_flag = fhd >> 6, (fhd >> 5) & 1, (fhd >> 2) & 1, fhd & 3
    if checksum or dict_flag or (fhd & 0x18):
        return None
    pos = 5 + (0 if single else 1)
    fcs_bytes = (1 if single else 0, 2, 4, 8)[fcs_flag]
    if fcs_bytes == 0 or pos + fcs_bytes > n:
        return None
    content = int.from_bytes(bytes(d[pos:pos + fcs_bytes]), "little") + (256 if fcs_bytes == 2 else 0)
    pos += fcs_bytes
    tables, table_ids, dst = [], {}, 0
    c_src, c_dst, c_type, c_regen, c_s0, c_s1, c_s2, c_s3, c_tab = [], [], [], [], [], [], [], [], []
    u32 = struct.Struct("<I").unpack_from
    jump = struct.Struct("<HHH").unpack_from
    last_tree, last_tid = None, 0
    while True:
        if pos + 3 > n:
            return None
        h = u32(d, pos)[0] & 0xFFFFFF if pos + 4 <= n else int(d[pos]) | int(d[pos + 1]) << 8 | int(d[pos + 2]) << 16
        last, btype, size = h & 1, (h >> 1) & 3, h >> 3
        pos += 3
        s0 = s1 = s2 = s3 = tid = 0
        if btype == 0:
            if pos + size > n:
                return None
            src, regen, pos = pos, size, pos + size
        elif btype == 1:
            if pos + 1 > n:
                return None
            src, regen, pos = pos, size, pos + 1
        elif btype == 2:
            end = pos + size
            if end > n or size < 16:
                return None
            b0 = d[pos]
            fmt = (b0 >> 2) & 3
            if b0 & 3 != 2 or fmt == 0:
                return None                                  # raw / RLE / treeless literals or a single stream
            lh = 2 + fmt
            v = (u32(d, pos)[0] | d[pos + 4] << 32) >> 4
            nbits = 6 + 4 * fmt                              # 10, 14 or 18 bits per size
            regen, csize = v & ((1 << nbits) - 1), (v >> nbits) & ((1 << nbits) - 1)
            if pos + lh + csize + 1 != end or d[end - 1] != 0 or regen < 6 or regen > BLOCK:
                return None                                  # sequences follow the literals
            head = d[pos + lh]
            used = 1 + ((head - 126) // 2 if head >= 128 else head)      # bytes of the tree description
            if used + 6 >= csize:
                return None
            tree = d[pos + lh:pos + lh + used]
            if tree != last_tree:                            # the blocks of a frame nearly always share one tree
                key = bytes(tree)
                try:
                    if key not in table_ids:
                        table_ids[key] = len(tables)
                        tables.append(decode_table(key)[0])
                except (ValueError, IndexError):
                    return None
                last_tree, last_tid = key, table_ids[key]
            tid = last_tid
            jt = pos + lh + used
            s0, s1, s2 = jump(d, jt)
            s3 = csize - used - 6 - s0 - s1 - s2
            if s0 == 0 or s1 == 0 or s2 == 0 or s3 <= 0:
                return None
            src, pos = jt + 6, end
        else:
            return None
        c_src.append(src); c_dst.append(dst); c_type.append(btype); c_regen.append(regen); c_tab.append(tid)
        c_s0.append(s0); c_s1.append(s1); c_s2.append(s2); c_s3.append(s3)
        dst += regen
        if last:
            break
    if pos != n or dst != content:
        return None
    blocks = np.zeros(len(c_src), DBLOCK)
    blocks["src_off"], blocks["dst_off"], blocks["type"], blocks["regen"], blocks["table"] = c_src, c_dst, c_type, c_regen, c_tab
    blocks["stream_bytes"] = np.array([c_s0, c_s1, c_s2, c_s3], np.uint32).T
    return content, blocks, (np.stack(tables) if tables else np.zeros((0, 1 << DLOG), np.uint16))


def decompress_device(data, device, max_bytes=None, marks=None):
    """data: bytes / u8 array with one zstd frame -> u8 CUDA tensor with its content, decoded by the kernels of
    csrc/tz_zstd.cu; None when the frame uses parts of the format they do not cover (decode it with libzstd then).
    The upload of a frame that carries this writer's header is queued before the host walks the block headers (from
    pinned memory the two overlap).  max_bytes: refuse frames that declare more content than this.  marks: optional
    list; CUDA events around the decoding kernels are appended as ("decode", start, end)."""
    import torch
    from . import _lib
    from .ops import check, ptr, _st
    arr = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    dev = torch.device(device)

    def upload():
        # (the stream decoder reads aligned 32-bit words: room for the word that holds the frame's last byte)
        f = torch.empty(arr.size + 8, dtype=torch.uint8, device=dev)
        f[:arr.size].copy_(torch.from_numpy(arr if arr.flags.writeable else arr.copy()), non_blocking=True)
        return f

    frame = upload() if arr.size > 6 and bytes(arr[4:6]) == b"\xc0\x38" else None
    parsed = parse_frame(arr)
    if parsed is None:
        return None
    content, blocks, tables = parsed
    if max_bytes is not None and content > max_bytes:
        raise RuntimeError("zstd frame declares %d bytes of content, more than the limit of %d "
                           "(TEZIP_MAX_DECODED_BYTES)" % (content, max_bytes))
    out = torch.empty(content, dtype=torch.uint8, device=dev)
    if content == 0:
        return out
    if frame is None:
        frame = upload()
    blk = torch.from_numpy(blocks.view(np.uint8).reshape(-1).copy()).to(dev)
    tab = torch.from_numpy(tables.view(np.int16).reshape(-1).copy()).to(dev) if len(tables) else None
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    cur = torch.cuda.current_stream(dev)
    e0 = cur.record_event(torch.cuda.Event(enable_timing=True)) if marks is not None else None
    check(_lib.load().tz_zstd_decode(ptr(frame), ptr(blk), len(blocks), ptr(tab), ptr(out), ptr(err), _st(dev)),
          "tz_zstd_decode")
    if marks is not None:
        marks.append(("decode", e0, cur.record_event(torch.cuda.Event(enable_timing=True))))
    code = int(err.item())
    if code:
        raise RuntimeError("corrupt zstd frame (Huffman stream error %d)" % code)
    return out


def frame_device(t, marks=None):
    """t: a contiguous CUDA tensor (any element type; its bytes are compressed) -> (u8 CUDA tensor, size): the first
    `size` bytes of the tensor are one zstd frame with content size.  One small device->host read in the middle (the
    1 KB histogram: the Huffman code is built on the host) and one at the end (the size).
    marks: optional list; CUDA events around the two device phases are appended as (name, start, end)."""
    import torch
    from . import _lib
    from .ops import check, ptr, _st
    if not t.is_cuda:
        raise ValueError("the zstd frame writer takes a CUDA tensor (there is no CPU path)")
    raw = t.contiguous().view(-1).view(torch.uint8)
    n = raw.numel()
    dev = raw.device
    if n == 0:
        return torch.from_numpy(np.frombuffer(empty_frame(), np.uint8).copy()).to(dev), len(empty_frame())
    lib = _lib.load()
    nblocks = -(-n // BLOCK)
    hist = torch.empty(256, dtype=torch.int32, device=dev)
    uniform = torch.empty(nblocks, dtype=torch.int32, device=dev)
    cur = torch.cuda.current_stream(dev)
    e0 = cur.record_event(torch.cuda.Event(enable_timing=True)) if marks is not None else None
    check(lib.tz_zstd_hist(ptr(raw), n, ptr(hist), ptr(uniform), _st(dev)), "tz_zstd_hist")
    if marks is not None:
        marks.append(("hist", e0, cur.record_event(torch.cuda.Event(enable_timing=True))))
    ct, tree = huffman_tables(hist.cpu().numpy().view(np.uint32))
    ct_dev = torch.from_numpy(ct.view(np.int32)).to(dev)
    tree_dev = torch.from_numpy(np.frombuffer(tree + b"\0", np.uint8).copy()).to(dev)
    ws = torch.empty(int(lib.tz_zstd_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    out = torch.empty(int(lib.tz_zstd_bound(n)), dtype=torch.uint8, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    e0 = cur.record_event(torch.cuda.Event(enable_timing=True)) if marks is not None else None
    check(lib.tz_zstd_encode(ptr(raw), n, ptr(ct_dev), ptr(tree_dev), len(tree), ptr(uniform), ptr(ws), ptr(out),
                             ptr(total), _st(dev)), "tz_zstd_encode")
    if marks is not None:
        marks.append(("encode", e0, cur.record_event(torch.cuda.Event(enable_timing=True))))
    return out, int(total.item())


_PINNED = {}


def frame_host(t):
    """-> u8 numpy view of the frame of frame_device(t) in a cached pinned host buffer (one per device, grown on
    demand): valid until the next call for that device.  For callers that write the frame out at once."""
    import torch
    out, size = frame_device(t)
    key = str(out.device)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < size:
        buf = _PINNED[key] = torch.empty(max(size, 1 << 20), dtype=torch.uint8, pin_memory=True)
    buf[:size].copy_(out[:size], non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return buf.numpy()[:size]


def compress_device(t):
    """-> bytes: the frame of frame_device(t), copied to the host."""
    return frame_host(t).tobytes()
