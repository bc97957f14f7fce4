#!/bin/bash
# 2-GPU checks: NCCL parity test of the view-sharded step, bench.py at N=2 (weak scaling + 8-view step)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu > gpurun_out/r02n2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n2_pytest.log
tail -12 gpurun_out/r02n2_pytest.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 \
    > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 exit $?"
grep "\[bench\]" gpurun_out/r02_bench_n2.err | tail -12 | cut -c1-300
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02_bench_n2.json'))
    print("N=2 value", d["value"], "ms", d["ms_per_step"], "phases", d.get("phases"), "launch", d["config"]["launch"])
    print("train_8_views", json.dumps(d.get("train_8_views"))[:600])
except Exception as e:
    print("no json:", e)
PY
