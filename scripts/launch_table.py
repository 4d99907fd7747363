"""Per-kernel table (ms/step, DRAM GB/step, GB/s) from an ncu csv with gpu__time_duration + dram bytes metrics."""
import csv, re, sys
from collections import OrderedDict

def main(path, launches_per_step):
    lines = [l for l in open(path) if not l.startswith('==')]
    per = {}
    for r in csv.DictReader(lines):
        k = r['ID']; per.setdefault(k, {'name': re.sub(r'\(.*', '', r['Kernel Name'])})
        scale = {'ns': 1, 'us': 1e3, 'ms': 1e6, 'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r['Metric Unit'], 1)
        per[k][r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * scale
    agg = OrderedDict()
    for v in per.values():
        a = agg.setdefault(v['name'], [0, 0.0, 0.0])
        a[0] += 1; a[1] += v.get('gpu__time_duration.sum', 0)
        a[2] += v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)
    tot = sum(a[1] for a in agg.values())
    steps = len(per) / float(launches_per_step)
    print(f"{len(per)} launches = {steps:.2f} steps; sum of kernel times {tot / 1e6 / steps:.2f} ms/step (ncu: cold cache, serialised)\n")
    print("| kernel | launches/step | ms/step | share | DRAM GB/step | GB/s |")
    print("|---|---:|---:|---:|---:|---:|")
    for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n[:70]}` | {c / steps:.1f} | {t / 1e6 / steps:.3f} | {100 * t / tot:.1f}% | {b / 1e9 / steps:.2f} | {b / max(t, 1):.0f} |")

if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1)
