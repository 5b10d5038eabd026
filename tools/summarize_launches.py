"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [passes] > profiles/<name>.md"""
import collections
import csv
import io
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        t *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row.get("Metric Unit", "ns"), 1.0)
        name = row["Kernel Name"]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    return agg


def main():
    agg = load(sys.argv[1])
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | avg us | max us | share |")
    print("|---|---:|---:|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        short = n.split("(")[0][-70:]
        print(f"| `{short}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {a[2] / 1e3:.1f} | {a[1] / tot:.4f} |")
    print(f"\ntotal device time {tot / 1e6:.2f} ms over {passes} forward pass(es) -> {tot / 1e6 / passes:.2f} ms per pass "
          "(ncu-serialised, cold cache: compare SHARES, not absolutes)")


if __name__ == "__main__":
    main()
