#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_loss_train.py -x -q -m gpu > gpurun_out/r02n2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n2b_pytest.log
tail -6 gpurun_out/r02n2b_pytest.log | cut -c1-400
bash scripts/gpu_run_n8.sh 2
