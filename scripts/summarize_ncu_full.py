"""Per-launch table from `ncu --set full` (ncu -i X.ncu-rep --page raw --csv > raw.csv): time, DRAM traffic, pipe use."""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA bytes"),
        ("launch__registers_per_thread", "regs"), ("smsp__warps_active.avg.per_cycle_active", "warps/SMSP")]


def main(raw, title, out_json=None):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("| # | kernel | grid x block | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    traffic = {}
    for n, r in enumerate(rows[2:]):
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in COLS:
            if key in ix and r[ix[key]] not in ("", "n/a"):
                v = r[ix[key]].replace(",", "")
                try:
                    cells.append("%.4g %s" % (float(v), units[ix[key]]))
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        print("| %d | `%s` | %s x %s | %s |" % (n, name, r[ix["Grid Size"]], r[ix["Block Size"]], " | ".join(cells)))

        def num(k):
            v, u = float(r[ix[k]].replace(",", "")), units[ix[k]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        traffic.setdefault(name, []).append({"grid": r[ix["Grid Size"]], "dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum")})
    if out_json:
        json.dump(traffic, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
