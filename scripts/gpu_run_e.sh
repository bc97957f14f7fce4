#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -25 gpurun_out/r02e_pytest.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_launches_train.csv python scripts/quick_train.py > gpurun_out/r02e_ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02e_launches_train.csv') if l.startswith('"')))
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
python scripts/quick_train.py 2>&1 | tail -2
( time python bench.py --steps 50 --warmup 5 --configs 1,2 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err ) 2>&1 | tail -3
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_n1.json')); print(json.dumps(d['configs'])[:3000]); print(d['value'], d['train'])"
