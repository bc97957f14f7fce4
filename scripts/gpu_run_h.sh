#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loss_train.py -x -q -m gpu 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches_train.csv python scripts/quick_train.py > gpurun_out/r02h_ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02h_launches_train.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1000.0
    elif r[ui] == 'ms': v *= 1000.0
    d[r[ki].split('(')[0]].append(v)
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:6]:
    print(f"{k[:70]:70s} n={len(v):4d} mean_us={sum(v)/len(v):9.1f}")
PY
python scripts/quick_train.py 2>&1 | tail -1
for b in 1 4; do timeout 30 scripts/dev/tma_gather4_probe $b; echo "probe box=$b rc=$?"; done
