#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep captured with
--import-source on (kernels are built with -lineinfo).  Usage: ncu_source_hotspots.py REPORT KERNEL_REGEX [TOP]"""
import csv
import subprocess
import sys


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
            continue
        try:
            ln, inst = int(r[0]), int(r[7])
            samp = int(r[4]) if r[4] not in ("-", "") else 0
        except ValueError:
            continue
        a = agg.setdefault((cur, ln), [0, 0, r[1][:100]])
        a[0] += inst
        a[1] += samp
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"kernel regex {regex}: {tot} warp instructions, {ts} stall samples")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{k[0]:>28s}:{k[1]:<4d} inst {a[0] / tot * 100:5.1f}%  samples {a[1] / ts * 100:5.1f}%  {a[2]}")


if __name__ == "__main__":
    main()
