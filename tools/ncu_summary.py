"""Tabulates `ncu --set full ... --page raw --csv` exports (profiles/r02_ncu_*_raw.csv) into a markdown summary.
usage: python tools/ncu_summary.py profiles/r02_ncu_vit_raw.csv [...] > profiles/r02_ncu_summary.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def table(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units = rows[hdr], rows[hdr + 1]
    cols = [(h.index(c), n, units[h.index(c)]) for c, n in COLS if c in h]
    kn = h.index("Kernel Name")
    out = ["| # | kernel | " + " | ".join(f"{n} ({u})" if u and u != "%" else n for _, n, u in cols) + " |",
           "|---:|---|" + "---:|" * len(cols)]
    for i, r in enumerate(rows[hdr + 2:]):
        if len(r) <= kn:
            continue
        name = r[kn].split("(")[0].replace("void ", "")[-64:]
        vals = []
        for c, _, _ in cols:
            try:
                v = float(r[c].replace(",", ""))
                vals.append(f"{v:.1f}" if abs(v) < 1000 else f"{v:.0f}")
            except ValueError:
                vals.append(r[c])
        out.append(f"| {i} | `{name}` | " + " | ".join(vals) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"### {p}\n")
        print(table(p))
        print()
