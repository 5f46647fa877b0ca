"""DRAM bytes per launch of the kernels of one next() from an `ncu --set full` raw page
(ncu -i X.ncu-rep --page raw --csv > raw.csv) -> profiles/r2_traffic.json (read by bench.py -> roofline.traffic).
The launches of one next() are named by their order: e0, a_0.., gates_{L-1}..gates_1, gates0_r1half, l0_tail."""
import csv
import json
import sys

ORDER = ["e0", "conv_tc_a0", "conv_tc_a1", "conv_tc_a2", "conv_tc_gates3", "conv_tc_gates2", "conv_tc_gates1",
         "conv_tc_gates0_r1half", "l0_tail"]


def main(raw, out, note):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        v, u = float(r[ix[k]].replace(",", "")), units[ix[k]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    seq = [r for r in rows[2:] if any(s in r[ix["Kernel Name"]] for s in ("e0_tc_kernel", "e0_px_kernel", "conv_tc_kernel", "l0_tail"))]
    start = next(i for i, r in enumerate(seq) if ("e0_tc_kernel" in r[ix["Kernel Name"]] or "e0_px_kernel" in r[ix["Kernel Name"]]))
    seq = seq[start:start + len(ORDER)]
    per = {}
    for name, r in zip(ORDER, seq):
        per[name] = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    json.dump({"source": note, "per_launch_dram_bytes": per}, open(out, "w"), indent=1)
    print(json.dumps(per, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
