"""Summarises an `ncu --set full` report as markdown rows (one per captured launch): duration, DRAM bytes, issue / warp
activity, tensor pipe, occupancy limits, top stall reasons, and -- with --phases -- executed instructions between the
kernel's barriers.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--phases]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("smsp__inst_executed.sum", "warp instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) %"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"), ("launch__waves_per_multiprocessor", "waves"),
        ("launch__grid_size", "grid"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, launches = raw(rep)
    name_i = hdr.index("Kernel Name")
    for L in launches:
        print("### `%s`" % L[name_i][:100])
        print()
        print("| metric | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in hdr:
                print("| %s | %s %s |" % (label, L[hdr.index(k)], units[hdr.index(k)]))
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(float(L[i].replace(",", "")))) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and L[i] not in ("", "n/a")]
        tot = sum(v for _k, v in st) or 1
        st.sort(key=lambda kv: -kv[1])
        print("| warp-state samples | " + ", ".join("%s %.0f%%" % (k, 100.0 * v / tot) for k, v in st[:7]) + " |")
        print()
    if "--phases" in sys.argv:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        iA, iE, iS = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        sass = [(r[iA].strip(), int(r[iE]), int(r[iS])) for r in rows[2:] if len(r) > iE and r[0].startswith("0x")]
        total, ts = sum(e for _s, e, _ in sass), sum(s for _s, _e, s in sass) or 1
        print("Executed warp instructions between consecutive `BAR.SYNC`s (segments above 1 %%), total %.1f M:" % (total / 1e6))
        print()
        print("| SASS range | warp instr (M) | share | stall-sample share |")
        print("|---|---:|---:|---:|")
        acc = sacc = start = 0
        for k, (src, ex, sm) in enumerate(sass):
            acc += ex
            sacc += sm
            if "BAR.SYNC" in src or k == len(sass) - 1:
                if acc > 0.01 * total:
                    print("| %d-%d | %.1f | %.1f %% | %.1f %% |" % (start, k, acc / 1e6, 100.0 * acc / total, 100.0 * sacc / ts))
                acc = sacc = 0
                start = k + 1


if __name__ == "__main__":
    main()
