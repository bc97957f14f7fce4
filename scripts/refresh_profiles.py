#!/usr/bin/env python
"""Regenerate the round-2 files under profiles/ from the scratch captures in gpurun_out/ (run here, after a gpurun capture):
   refresh_profiles.py FULL_REPORT.ncu-rep FWD_BWD_LAUNCHES.csv TRAIN_LAUNCHES.csv
FULL_REPORT: `ncu --set full --import-source on` of one training step (scripts/quick_train.py) -- every kernel of the step."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
TAG = "r02"


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def launch_table(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
        d.setdefault(r[ki].split("(")[0], []).append(v)
    return d


def main():
    rep, launches, train = sys.argv[1], sys.argv[2], sys.argv[3]
    shutil.copy(launches, os.path.join(P, f"{TAG}_launches.csv"))
    shutil.copy(train, os.path.join(P, f"{TAG}_launches_train_step.csv"))
    open(os.path.join(P, f"{TAG}_all_kernels_ncu_full.txt"), "w").write(run([sys.executable, os.path.join(ROOT, "scripts", "summarize_ncu.py"), rep]))
    hot = []
    for k in ("blend_bwd", "blend_fwd", "project_cull", "project_visible", "project_bwd", "onesweep", "image_loss_fused", "train_update"):
        hot.append(run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_hotspots.py"), rep, k, "16"]))
    open(os.path.join(P, f"{TAG}_source_hotspots.txt"), "w").write("\n".join(hot))
    open(os.path.join(P, f"{TAG}_sass_histograms.txt"), "w").write(run([sys.executable, os.path.join(ROOT, "scripts", "sass_histogram.py")]))
    # DRAM traffic of the dominant kernel, per launch
    raw = run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:blend_bwd", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"])
    rows = list(csv.reader(raw.splitlines()))
    hdr = next(r for r in rows if "Kernel Name" in r)
    units, vals = rows[rows.index(hdr) + 1], rows[rows.index(hdr) + 2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(vals[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    json.dump({"blend_bwd_kernel": tot, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
               "source": f"profiles/{TAG}_all_kernels_ncu_full.txt (ncu --set full, config [1] scene, one launch)"},
              open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    for name, path in (("fwd+bwd", launches), ("train step", train)):
        d = launch_table(path)
        total = sum(sum(v) / len(v) * (len(v) / max(len(next(iter(d.values()))), 1)) for v in d.values())
        print(f"-- {name}: per-launch mean us")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            print(f"   {k[:64]:64s} n={len(v):4d} mean={sum(v) / len(v):8.1f}")
    print("DRAM bytes per launch of blend_bwd_kernel:", tot)


if __name__ == "__main__":
    main()
