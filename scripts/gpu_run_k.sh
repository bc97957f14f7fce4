#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
tail -4 gpurun_out/r02k_pytest.log | cut -c1-300
( time timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | tail -3
grep "\[bench\]" gpurun_out/r02_bench_n1.err | tail -3
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_n1.json'))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launch", d["config"]["launch"])
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["pipeline_fwd_bwd"]["frac"])
print("train", d.get("train"))
print("configs", json.dumps(d.get("configs"))[:1800])
print("cpu", d.get("cpu_baseline"), d.get("clocks"))
PY
