"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`: instructions with the most
stall samples, with their dominant stall reason."""
import csv
import sys


def split(path):
    """The page holds one block per launch: a "Kernel Name" row, a header row, then the instructions."""
    blocks, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = [r]
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    return blocks


def main(path, top=40, which=0):
    rows = split(path)[which]
    print(rows[0][1])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = 0
    for n, r in enumerate(rows[2:]):
        if len(r) < len(hdr):
            continue
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
        data.append((s, n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), st))
    print("total samples", tot, "instructions", len(data))
    agg = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        for c in stall_cols:
            agg[c] = agg.get(c, 0) + int(r[ix[c]] or 0)
    print("stall totals:", ", ".join("%s %d" % (c[6:], v) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for s, n, src, ex, st in sorted(data, reverse=True)[:top]:
        print("%6d %5.1f%% #%5d ex %8d  %-70s %s" % (s, 100.0 * s / max(tot, 1), n, ex, src[:70],
                                                    " ".join("%s:%d" % (c[6:], v) for v, c in st if v)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
