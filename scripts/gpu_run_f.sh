#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -40 gpurun_out/r02f_pytest.log | cut -c1-300
python scripts/quick_train.py 2>&1 | tail -1
python scripts/train_colmap.py --iters 3000 --hidden 2000000 > gpurun_out/r02f_train_colmap.json 2> gpurun_out/r02f_train_colmap.err; tail -3 gpurun_out/r02f_train_colmap.err | cut -c1-1500
