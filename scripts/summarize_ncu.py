"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys


def main(path, title):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = row["Kernel Name"].split("(")[0]
        grid = row["Grid Size"]
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1
        a[1] += v
        a[2].add(grid)
    tot = sum(a[1] for a in agg.values())
    print("# %s\n" % title)
    print("Source: `%s` (ncu launch list, cold-cache and serialised: compare SHARES, not absolutes).\n" % path)
    print("| kernel | launches | total ms | avg ms | share |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.1f%% |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    print("\ntotal %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
