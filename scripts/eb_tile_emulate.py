"""Lane-level emulation (numpy, 32-wide) of the tiled error-bound kernel in tezip_b200/csrc/tz_codec.cu
(eb_plane_tiles), checked against the oracle's restatement of compress.py:23-70.  Development aid: it lets the
chunk/tile logic be verified on the CPU before spending GPU time.  Usage: python scripts/eb_tile_emulate.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import codec_oracle as co  # noqa: E402

ID = (32767, -32768)   # identity of join: (min, max)
LANES = np.arange(32)


def join(p, q):
    return (np.minimum(p[0], q[0]), np.maximum(p[1], q[1]))


def broken(p, G):
    return (p[1].astype(np.int64) - p[0].astype(np.int64)) > G


def shfl(p, src):
    src = np.asarray(src) & 31
    return (np.asarray(p[0])[src], np.asarray(p[1])[src])


def mid(st, E):
    a, b = float(st[0]), float(st[1])
    return np.int16(int(((a + E) + (b - E)) * 0.5))   # trunc toward zero like the float -> int slice assignment


def eb_tiles(d, E, G, T=64):
    """d: int16 1-D plane (copy is modified and returned)."""
    d = d.astype(np.int16).copy()
    n = len(d)
    nchT = T // 32
    carry = (np.int64(ID[0]), np.int64(ID[1]))
    head = 0
    for tb in range(0, n, T):
        Tn = min(T, n - tb)
        nch = (Tn + 31) // 32
        spre = [None] * nchT
        srun = [None] * nchT
        snb = np.zeros((nchT, 32), np.int64)
        slast = np.zeros((nchT, 32), np.int64)
        sentry = np.full(nchT, 255)
        sinst = [None] * nchT
        soutst = [None] * nchT
        # ---- S1
        for c in range(nch):
            i = 32 * c + LANES
            inn = i < Tn
            dv = np.where(inn, d[np.minimum(tb + i, n - 1)], 0).astype(np.int64)
            v = (np.where(inn, dv, ID[0]), np.where(inn, dv, ID[1]))
            pre = v
            off = 1
            while off < 32:
                t = shfl(pre, np.maximum(LANES - off, 0))
                jn = join(pre, t)
                pre = (np.where(LANES >= off, jn[0], pre[0]), np.where(LANES >= off, jn[1], pre[1]))
                off <<= 1
            M = [v]
            for k in range(4):
                o = 1 << k
                t = shfl(M[k], LANES + o)
                t = (np.where(LANES + o < 32, t[0], ID[0]), np.where(LANES + o < 32, t[1], ID[1]))
                M.append(join(M[k], t))
            run = v
            pos = LANES + 1
            for k in range(4, -1, -1):
                cand = shfl(M[k], pos)
                nx = join(run, cand)
                ok = (pos + (1 << k) <= 32) & ~broken(nx, G)
                run = (np.where(ok, nx[0], run[0]), np.where(ok, nx[1], run[1]))
                pos = np.where(ok, pos + (1 << k), pos)
            e = pos.copy()
            l = LANES.copy()
            for _r in range(5):
                te = e[e & 31]
                tl = l[e & 31]
                go = e < 32
                e = np.where(go, te, e)
                l = np.where(go, tl, l)
            assert (e >= 32).all()
            spre[c], srun[c], snb[c], slast[c] = pre, run, pos, l
        # ---- S2 (warp 0): the serial chain records (entry, X on entry, previous closing chunk) per closing chunk
        X = carry
        hd = head
        seg_chunk = -1
        wb = None
        sinx = [None] * nchT
        sprev = [None] * nchT
        for c in range(nch):
            p = spre[c]
            t = join((np.full(32, X[0]), np.full(32, X[1])), p)
            m = broken(t, G)
            if not m.any():
                X = (t[0][31], t[1][31])
                continue
            b = int(np.argmax(m))
            sentry[c] = b
            sinx[c] = X
            sprev[c] = seg_chunk
            l = int(slast[c][b])
            X = (srun[c][0][l], srun[c][1][l])
            hd = tb + 32 * c + l
            seg_chunk = c
        # ---- S2b (lane c owns chunk c): closing states and the chunks they cover
        for c in range(nch):
            b = int(sentry[c])
            if b == 255:
                continue
            xin = sinx[c]
            Xc = join(xin, (spre[c][0][b - 1], spre[c][1][b - 1])) if b > 0 else xin
            pc = sprev[c]
            sinst[c] = Xc
            if pc >= 0:
                soutst[pc] = Xc
            else:
                wb = (head, Xc)
            for cc in range(pc + 1, c):
                sinst[cc] = Xc
        carry, head = X, hd
        open_from = hd - tb if hd >= tb else 0
        # ---- write-back of the part of a closed segment that lies in earlier tiles
        if wb is not None and wb[0] < tb:
            d[wb[0]:tb] = mid(wb[1], E)
        # ---- S3
        for c in range(nch):
            ent = int(sentry[c])
            nbrel = snb[c]
            R = 0
            if ent != 255:
                R = 1 << ent
                J = nbrel.copy()
                for _r in range(5):
                    tgt = 0
                    for j in range(32):
                        if (R >> j) & 1 and J[j] < 32:
                            tgt |= 1 << int(J[j])
                    R |= tgt
                    J = np.where(J < 32, J[J & 31], 32)
            for j in range(32):
                i = 32 * c + j
                if i >= Tn or i >= open_from:
                    continue
                if ent == 255 or j < ent:
                    st = sinst[c]
                else:
                    msk = R & (0xFFFFFFFF >> (31 - j))
                    s = msk.bit_length() - 1
                    if nbrel[s] >= 32:
                        st = soutst[c]
                    else:
                        st = (srun[c][0][s], srun[c][1][s])
                assert st is not None, (tb, c, j)
                d[tb + i] = mid(st, E)
    if head < n:
        d[head:n] = mid(carry, E)
    return d


def main():
    rng = np.random.default_rng(0)
    cases = 0
    for trial in range(400):
        n = int(rng.choice([1, 5, 31, 32, 33, 63, 64, 65, 100, 128, 200, 333, 700]))
        amp = int(rng.choice([0, 1, 3, 8, 40, 300]))
        d = rng.integers(-amp, amp + 1, size=n).astype(np.int64)
        if trial % 7 == 0:   # long flat stretches with rare spikes
            d = np.zeros(n, np.int64)
            d[rng.integers(0, n, size=max(1, n // 50))] = rng.integers(-30, 30)
        if trial % 11 == 0:
            d = np.cumsum(rng.integers(-1, 2, size=n))
        E = float(rng.choice([0.5, 1.0, 2.0, 2.5, 7.0]))
        G = int(np.floor(2 * E))
        T = int(rng.choice([32, 64, 128]))
        orig = np.zeros(n, np.int64)
        want = co.error_bound_py(orig, d.copy().astype(np.float64), "abs", [E]).astype(np.int64)
        got = eb_tiles(d, E, G, T).astype(np.int64)
        if not np.array_equal(want, got):
            bad = np.flatnonzero(want != got)
            print("MISMATCH trial", trial, "n", n, "T", T, "E", E, "first bad", bad[:5], want[bad[:5]], got[bad[:5]])
            return 1
        cases += 1
    print("ok", cases, "cases")
    return 0


if __name__ == "__main__":
    sys.exit(main())
