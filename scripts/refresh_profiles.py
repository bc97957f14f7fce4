#!/usr/bin/env python
"""Regenerate profiles/ from the scratch captures in gpurun_out/ (run here, after a gpurun capture):
   refresh_profiles.py REPORT.ncu-rep LAUNCHES.csv [TRAIN_LAUNCHES.csv]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep, launches = sys.argv[1], sys.argv[2]
    shutil.copy(launches, os.path.join(P, "r01_launches.csv"))
    if len(sys.argv) > 3:
        shutil.copy(sys.argv[3], os.path.join(P, "r01_launches_train_step.csv"))
    summ = run([sys.executable, os.path.join(ROOT, "scripts", "summarize_ncu.py"), rep])
    open(os.path.join(P, "r01_all_kernels_ncu_full.txt"), "w").write(summ)
    hot = []
    for k in ("rasterize_bwd", "rasterize_fwd", "project_cull", "project_visible", "project_bwd", "onesweep"):
        hot.append(run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_hotspots.py"), rep, k, "18"]))
    open(os.path.join(P, "r01_source_hotspots.txt"), "w").write("\n".join(hot))
    # DRAM traffic of the dominant kernel
    raw = run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:rasterize_bwd",
               "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"])
    rows = list(csv.reader(raw.splitlines()))
    hdr = next(r for r in rows if "Kernel Name" in r)
    vals = rows[rows.index(hdr) + 2]
    units = rows[rows.index(hdr) + 1]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        tot += float(vals[i].replace(",", "")) * mult
    json.dump({"rasterize_bwd_kernel": int(tot), "source": os.path.basename(rep)}, open(os.path.join(P, "ncu_traffic.json"), "w"))
    print("traffic", int(tot))


if __name__ == "__main__":
    main()
