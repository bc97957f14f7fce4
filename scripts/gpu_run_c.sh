#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_loss_train.py -x -q -m gpu > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -30 gpurun_out/r02c_pytest.log
python scripts/quick_time.py > gpurun_out/r02c_quick.log 2>&1; tail -3 gpurun_out/r02c_quick.log
