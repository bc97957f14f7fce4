#!/bin/bash
# Round-2 evidence run (one GPU): GPU tests, the default bench line, launch lists and one ncu --set full capture of a training step.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/r02l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest.log
tail -3 gpurun_out/r02l_pytest.log | cut -c1-300
( time timeout 500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_n1.json'))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("roofline", json.dumps(d["roofline"])[:600])
print("train", d.get("train"))
print("configs", json.dumps(d.get("configs"))[:2500])
print("cpu", d.get("cpu_baseline"), d.get("clocks"))
PY
# launch lists (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --configs 1 --no-cpu-baseline --no-train > gpurun_out/r02l_ncu_bench.log 2>&1; echo "ncu bench rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv \
    python scripts/quick_train.py > gpurun_out/r02l_ncu_train.log 2>&1; echo "ncu train rc $?"
# every kernel of one training step, full sections with source
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name regex:"^(blend_|project_|onesweep_|gather_scan|tile_offsets|image_loss|train_update|bump_epoch)" --launch-skip 96 --launch-count 16 \
    -o gpurun_out/r02_train_step_full -f python scripts/quick_train.py > gpurun_out/r02l_ncu_full.log 2>&1; echo "ncu full rc $?"
ls -la gpurun_out/r02_train_step_full.ncu-rep
python scripts/quick_train.py 2>&1 | tail -1
