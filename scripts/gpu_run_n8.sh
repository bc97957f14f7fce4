#!/bin/bash
# 8-GPU bench (weak scaling of fwd+bwd with the exchange, and the 8-view step at 2M Gaussians)
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 30 --warmup 5 \
    > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N exit $?"
grep "\[bench\]" gpurun_out/r02_bench_n$N.err | tail -4 | cut -c1-200
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02_bench_n$N.json'))
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "phases", d.get("phases"))
    print("train_8_views", json.dumps(d.get("train_8_views"))[:700])
except Exception as e:
    print("no json:", e)
PY
