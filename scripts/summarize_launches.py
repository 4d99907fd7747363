"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and share.
Usage: python scripts/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/launches_rNN.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, steps=1):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    agg = OrderedDict()
    for name, ns, g, b in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.1f} | {100 * ns / total:.1f}% |")
    print(f"| **total** | {len(rows)} | {total / 1e6:.3f} | | 100% |")
    print(f"\n{len(rows)} launches captured ({steps} step(s)); ncu times are cold-cache and serialised: compare shares, not absolutes.")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
