#!/bin/bash
# N ranks: NCCL tests, the bench (config [4] leg), then the same with the device timeline of the multi-device step (BG_DP_TRACE)
mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu 2>&1 | tail -2
run() {  # $1 = tag, env passes through
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 \
        --configs 4 --no-cpu-baseline > gpurun_out/r02_$1_n$N.json 2> gpurun_out/r02_$1_n$N.err; echo "$1 n$N exit $?"
    python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02_$1_n$N.json'))
    print("value", d["value"], d.get("phases"))
    print("train_8_views", json.dumps(d.get("train_8_views"))[:400])
except Exception as e:
    print("no json:", e)
PY
}
run clean
BG_DP_TRACE=1 run trace
grep "bg dp trace" gpurun_out/r02_trace_n$N.err | tail -10
