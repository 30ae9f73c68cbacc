#!/usr/bin/env python
"""Per-kernel summary of one step out of an ncu launch list (tools/r02_launches.sh):  python tools/launch_summary.py FILE [step]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lines = [l for l in open(path) if l.startswith('"')]
    per = {}
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        e = per.setdefault(k, {"name": r["Kernel Name"]})
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        e["unit_" + r["Metric Name"]] = r["Metric Unit"]
    L = [per[i] for i in sorted(per)]
    names = [x["name"].split("(")[0].split("<")[0].replace("void ", "").replace("kdnb::", "") for x in L]
    kicks = [i for i, nm in enumerate(names) if nm == "kick_drift_kernel"]
    a, b = (kicks[which - 1] + 1 if which > 0 else 0), kicks[which] + 1
    agg = collections.OrderedDict()
    for i in range(a, b):
        x = L[i]
        t = x["gpu__time_duration.sum"] * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[x["unit_gpu__time_duration.sum"]]
        by = lambda k: x.get(k, 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[x.get("unit_" + k, "byte")]
        d = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
        e = agg.setdefault(names[i], [0, 0.0, 0.0])
        e[0] += 1
        e[1] += t
        e[2] += d
    tot = sum(v[1] for v in agg.values())
    print(f"{path}: step {which}: {b - a} launches, {tot:.1f} us under ncu (cold, serialised)")
    print("| kernel | launches | us/launch | us/step | share | DRAM MB/launch | TB/s |")
    print("|---|---|---|---|---|---|---|")
    for nm, (c, t, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{nm}` | {c} | {t / c:.1f} | {t:.1f} | {t / tot:.3f} | {d / c / 1e6:.1f} | {d / (t * 1e-6) / 1e12 if t else 0:.2f} |")


if __name__ == "__main__":
    main()
