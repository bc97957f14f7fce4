#!/bin/bash
# round-2 GPU check A: parity suite, blend A/B (per-lane masks vs vote masks), launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python scripts/quick_time.py > gpurun_out/r02a_quick_novote.log 2>&1; tail -4 gpurun_out/r02a_quick_novote.log
BG_BLEND_VOTE=1 python scripts/quick_time.py > gpurun_out/r02a_quick_vote.log 2>&1; tail -2 gpurun_out/r02a_quick_vote.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_launches.csv python scripts/quick_time.py > gpurun_out/r02a_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02a_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1000.0
    elif r[ui] == 'ms': v *= 1000.0
    d[r[ki].split('(')[0]].append(v)
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:70]:70s} n={len(v):4d} mean_us={sum(v)/len(v):9.1f}")
PY
