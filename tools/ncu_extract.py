#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the walk kernel into profiles/ncu_walk.json (read by bench.py for
roofline.traffic / roofline.ncu — bench.py never carries such numbers as literals) and export the raw / per-SASS pages.

    python tools/ncu_extract.py gpurun_out/walk_TAG.ncu-rep --key "n=1000000,gpus=1" --tag r02_walk_1M
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def page(rep, name, extra=()):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--key", required=True)
    ap.add_argument("--tag", required=True)
    a = ap.parse_args()
    raw = page(a.rep, "raw")
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    get = lambda k: float(vals[hdr.index(k)].replace(",", ""))
    unit = lambda k: rows[1][hdr.index(k)]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = get("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + get("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {
        "capture": f"profiles/{a.tag}_raw.csv", "commit": commit, "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "walk2_kernel",
        "duration_ms": get("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit("gpu__time_duration.sum")],
        "dram_bytes": dram,
        "fp64_pipe_active_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": get("smsp__inst_executed.sum"),
        "registers_per_thread": get("launch__registers_per_thread"),
        "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
    }
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{a.tag}_raw.csv"), "w").write(raw)
    open(os.path.join(ROOT, "profiles", f"{a.tag}_sass.csv"), "w").write(page(a.rep, "source", ["--print-source", "sass"]))
    path = os.path.join(ROOT, "profiles", "ncu_walk.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[a.key] = rec
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
