"""Per-kernel table (ms/step, DRAM GB/step, GB/s, tensor-pipe activity) from an ncu csv of
`bench.py --steps 2 --warmup 3 --kernels-only` captured with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv
The last `steps` training steps are cut out by their closing adam_step launch.  Also writes the DRAM-bytes-per-launch record
bench.py reports as roofline.traffic.

  python scripts/launch_table.py gpurun_out/r2_launches.csv [steps] [traffic.json] > profiles/r2_launches.md
"""
import csv
import json
import re
import sys
from collections import OrderedDict

TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"


def main(path, steps=2, traffic_out=None):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    per = OrderedDict()
    for r in csv.DictReader(lines):
        k = r["ID"]
        per.setdefault(k, {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
        scale = {"ns": 1, "us": 1e3, "ms": 1e6, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
        per[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * scale
    L = list(per.values())
    ends = [i for i, v in enumerate(L) if "adam_step" in v["name"] or "sgd_step" in v["name"]]
    S = L[ends[-steps - 1] + 1:ends[-1] + 1] if len(ends) > steps else L
    agg = OrderedDict()
    for v in S:
        a = agg.setdefault(v["name"], [0, 0.0, 0.0, 0.0])
        t = v.get("gpu__time_duration.sum", 0)
        a[0] += 1
        a[1] += t
        a[2] += v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)
        a[3] += t * v.get(TP, 0)
    tot = sum(a[1] for a in agg.values())
    print(f"{len(S)} launches = {steps} steps; sum of kernel times {tot / 1e6 / steps:.2f} ms/step (ncu: serialised, cold cache, "
          "unthrottled clocks: compare shares, not absolutes)\n")
    print("| kernel | launches/step | ms/step | share | DRAM GB/step | GB/s | tensor pipe active (time-weighted) |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    conv = {"igemm": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    other = 0.0
    for n, (c, t, b, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fam = "igemm" if "conv_igemm" in n else ("wgrad" if "conv_wgrad" in n else None)
        tpc = f"{tp / t:.1f} %" if fam else ""
        print(f"| `{n[:78]}` | {c / steps:.1f} | {t / 1e6 / steps:.3f} | {100 * t / tot:.1f}% | {b / 1e9 / steps:.2f} | {b / max(t, 1):.0f} | {tpc} |")
        if fam:
            conv[fam][0] += t
            conv[fam][1] += tp
        else:
            other += t
    ct, ctp = conv["igemm"][0] + conv["wgrad"][0], conv["igemm"][1] + conv["wgrad"][1]
    if ct > 0:
        print(f"\nTensor-core kernels: {ct / 1e6 / steps:.2f} ms/step, tensor pipe active {ctp / ct:.1f} % time-weighted over all their "
              f"launches (conv_igemm {conv['igemm'][1] / max(conv['igemm'][0], 1):.1f} %, conv_wgrad {conv['wgrad'][1] / max(conv['wgrad'][0], 1):.1f} %).")
    print(f"Everything else: {other / 1e6 / steps:.2f} ms/step.")
    if traffic_out:
        def fam_rec(key):
            xs = [v for v in S if key in v["name"]]
            return {"launches_captured": len(xs),
                    "dram_bytes_per_launch": sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in xs) / len(xs),
                    "avg_launch_us_under_ncu": sum(v["gpu__time_duration.sum"] for v in xs) / len(xs) / 1e3}
        json.dump({"conv_igemm": fam_rec("conv_igemm"), "conv_wgrad": fam_rec("conv_wgrad"),
                   "source": f"{path}: ncu launch list of `bench.py --steps 2 --warmup 3 --kernels-only`, last {steps} steps"},
                  open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2, sys.argv[3] if len(sys.argv) > 3 else None)
