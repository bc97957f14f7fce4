#!/bin/bash
# 2-GPU check: NCCL exchange tests, the multi-view step tests, then the 2-GPU bench
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_loss_train.py -x -q -m gpu 2>&1 | tail -6
bash scripts/gpu_run_n8.sh 2
