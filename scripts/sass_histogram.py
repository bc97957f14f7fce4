#!/usr/bin/env python
"""SASS instruction histograms of the shipped kernels (cuobjdump -sass of brush_b200/libbrush_b200.so), for profiles/:
which kernels use TMA (UBLKCP / UTMALDG), packed FP32 (FFMA2 / FMUL2 / FADD2), MUFU, shuffles, RED atomics, LDGSTS.
  python scripts/sass_histogram.py > profiles/r02_sass_histograms.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "brush_b200", "libbrush_b200.so")
KEY = ["UTMALDG", "UBLKCP", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "FSEL", "FSETP", "FMNMX", "SHFL", "REDG", "RED",
       "ATOMG", "REDUX", "VOTE", "LDS", "STS", "LDG", "STG", "SYNCS", "BAR", "CALL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    only = sys.argv[1:]
    for fn, h in hist.items():
        d = demangle(fn)
        short = re.sub(r"\(.*", "", d)
        if only and not any(o in short for o in only):
            continue
        total = sum(h.values())
        picks = [(k, sum(v for op, v in h.items() if op == k or (k in ("RED",) and op.startswith("RED") and not op.startswith("REDUX") and op != "REDG")))
                 for k in KEY]
        line = " ".join(f"{k}={v}" for k, v in picks if v)
        print(f"{short}\n    {total} instructions: {line}")


if __name__ == "__main__":
    main()
