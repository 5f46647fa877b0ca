"""Lane-level emulation (numpy, 32 lanes in lock step) of the speculative error-bound scan of
tezip_b200/csrc/tz_codec_lossy.cu (eb_spec_tile), checked against the oracle's restatement of compress.py:23-70.
Development aid: the tile / round / deferred-write logic is verified on the CPU before spending GPU time.

    python scripts/eb_spec_emulate.py

The greedy scan of compress.py:55-67 is serial: a segment is flushed when the next element would empty the running
interval intersection (for a plane-wide integer test: max - min > G).  Each of the 32 lanes owns L consecutive
elements of a tile and runs the scan over them from a FRESH state (speculation).  A greedy scan re-synchronises with
the true one as soon as both break at the same element, so lane k only has to re-walk its range from the true state
that enters it until one of its own speculative breaks is hit; ranges without any break are absorbed in O(1) through
their (min, max) summary.  The re-walks run for all lanes at once from the neighbours' current end states and are
repeated while some end state still changes (one round in the common case, at most 31).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import codec_oracle as co  # noqa: E402

BIG = 1 << 30
LANES = 32


def mid(mn, mx, E, exact):
    if exact:
        s = int(mn) + int(mx)
        return int(abs(s) // 2) * (1 if s >= 0 else -1)          # C integer division truncates toward zero
    return int(((float(mn) + E) + (float(mx) - E)) * 0.5)          # compress.py:61, float -> int truncation


def eb_spec(d, E, G, exact, L=8, stats=None):
    d = np.asarray(d, np.int64).copy()
    n = len(d)
    T = LANES * L
    carry = (BIG, -BIG)          # state of the open segment entering the tile
    head = 0                     # index where that segment starts
    for t0 in range(0, n, T):
        cnt = np.clip(n - t0 - np.arange(LANES) * L, 0, L)        # valid elements per lane
        e = np.zeros((L, LANES), np.int64)
        for k in range(LANES):
            e[:cnt[k], k] = d[t0 + k * L:t0 + k * L + cnt[k]]

        def walk(k, st, old_flags=None, old_end=None):
            """greedy scan of lane k's range from state st.  -> (flags, end state, steps)."""
            mn, mx = st
            flags = 0
            for j in range(cnt[k]):
                v = e[j, k]
                nmn, nmx = min(mn, v), max(mx, v)
                if nmx - nmn > G:
                    if old_flags is not None and (old_flags >> j) & 1:     # both scans break here: synchronised
                        keep = old_flags & ~((1 << j) - 1)
                        return flags | keep, old_end, j + 1
                    flags |= 1 << j
                    mn = mx = v
                else:
                    mn, mx = nmn, nmx
            return flags, (mn, mx), cnt[k]

        # ---- phase 1: speculation
        flags = [0] * LANES
        end = [None] * LANES
        summ = [None] * LANES
        for k in range(LANES):
            st = carry if k == 0 else (BIG, -BIG)
            flags[k], end[k], _ = walk(k, st)
            summ[k] = (int(e[:cnt[k], k].min()), int(e[:cnt[k], k].max())) if cnt[k] else (BIG, -BIG)
        # ---- phase 2: correction rounds
        inst = [carry] + [None] * (LANES - 1)
        active = [False] + [True] * (LANES - 1)
        rounds = 0
        while any(active):
            rounds += 1
            new_end = list(end)
            steps = 0
            for k in range(1, LANES):
                if not active[k]:
                    continue
                st = end[k - 1]                    # all lanes read their neighbour's end state of the PREVIOUS round
                inst[k] = st
                j0 = (min(st[0], summ[k][0]), max(st[1], summ[k][1]))
                if j0[1] - j0[0] <= G:             # the whole range joins the entering segment
                    flags[k], new_end[k] = 0, j0
                else:
                    flags[k], new_end[k], s = walk(k, st, flags[k], end[k])
                    steps = max(steps, s)
            nxt = [False] * LANES
            for k in range(1, LANES - 1):
                if active[k] and new_end[k] != end[k]:
                    nxt[k + 1] = True
            end, active = new_end, nxt
            if stats is not None:
                stats.append(steps)
        if stats is not None:
            stats.append(-rounds)
        # ---- phase 3: values.  forward: the value of every segment that closes inside the range is parked on its
        # last element; firstclose = value of the segment that enters the range, if it closes here.
        firstclose = [None] * LANES
        for k in range(LANES):
            mn, mx = inst[k]
            for j in range(cnt[k]):
                v = e[j, k]
                if (flags[k] >> j) & 1:
                    q = mid(mn, mx, E, exact)
                    if firstclose[k] is None:
                        firstclose[k] = q
                    if j >= 1:
                        e[j - 1, k] = q
                    mn = mx = v
                else:
                    mn, mx = min(mn, v), max(mx, v)
        tail = [None] * LANES          # value of the segment open at the END of lane k's range (closed by a later lane)
        nxt = None
        for k in range(LANES - 1, -1, -1):
            tail[k] = nxt
            if firstclose[k] is not None:
                nxt = firstclose[k]
        first_q = nxt                  # the carried segment closes in this tile with this value (None: still open)
        if first_q is not None and head < t0:
            d[head:t0] = first_q       # deferred elements of earlier tiles
        open_from = None
        for k in range(LANES):
            cur = tail[k]
            for j in range(cnt[k] - 1, -1, -1):
                if j + 1 < cnt[k] and (flags[k] >> (j + 1)) & 1:
                    cur = e[j, k]
                if cur is not None:
                    d[t0 + k * L + j] = cur
        # ---- carry
        last_break = None
        for k in range(LANES):
            if flags[k]:
                last_break = t0 + k * L + (flags[k].bit_length() - 1)
        if last_break is not None:
            head = last_break
        carry = end[max(k for k in range(LANES) if cnt[k] > 0)]
    q = mid(carry[0], carry[1], E, exact)                          # compress.py:67
    d[head:n] = q
    return d


def oracle_plane(d, E):
    ref = np.asarray(d, np.int64).reshape(1, -1, 1, 1).repeat(2, axis=0).copy()     # frame 0 is skipped by the oracle
    co.error_bound_frames(np.zeros_like(ref), ref, "abs", [E])
    return ref[1].ravel()


def main():
    rng = np.random.default_rng(1)
    cases = 0
    for L in (2, 8, 64):
        for n in (1, 5, 63, 64, 65, 257, 700, 2048 + 17, 5000):
            for kind in ("noise", "wide", "flat", "walk", "const", "ramp", "extremes"):
                for E in (0.5, 1.0, 2.0, 2.55, 7.9, 40.0):
                    if kind == "noise":
                        d = rng.integers(-3, 4, n)
                    elif kind == "wide":
                        d = rng.integers(-300, 301, n)
                    elif kind == "flat":
                        d = np.zeros(n, np.int64)
                        d[rng.integers(0, n, max(1, n // 300))] = rng.integers(-30, 31, max(1, n // 300))
                    elif kind == "walk":
                        d = np.cumsum(rng.integers(-1, 2, n))
                    elif kind == "const":
                        d = np.full(n, 5)
                    elif kind == "ramp":
                        d = (np.arange(n) // 37) - 20
                    else:
                        d = rng.choice(np.array([-32768, -32767, -1, 0, 1, 32766, 32767]), n)
                    twoE = 2 * E
                    G = int(np.floor(twoE))
                    exact = (E * 2 ** 36) == np.floor(E * 2 ** 36) and E < 4096
                    st = []
                    got = eb_spec(d, E, G, exact, L, st)
                    ref = oracle_plane(d, E)
                    if not np.array_equal(got, ref):
                        bad = np.nonzero(got != ref)[0]
                        raise SystemExit("MISMATCH L=%d n=%d %s E=%g at %s: got %s ref %s" % (
                            L, n, kind, E, bad[:5], got[bad[:5]], ref[bad[:5]]))
                    cases += 1
    print("ok:", cases, "cases")
    # cost model on BASELINE-like data: 128x160 plane, noise of sigma 2 around a smooth pattern, abs 2
    d = np.rint(rng.normal(0, 2.0, 20480)).astype(np.int64)
    st = []
    eb_spec(d, 2.0, 4, True, 64, st)
    rounds = [-s for s in st if s < 0]
    steps = [s for s in st if s >= 0]
    print("noise sigma 2, G=4, L=64: tiles %d, rounds per tile mean %.2f max %d, correction steps per round mean %.1f max %d"
          % (len(rounds), np.mean(rounds), max(rounds), np.mean(steps), max(steps)))


if __name__ == "__main__":
    main()
