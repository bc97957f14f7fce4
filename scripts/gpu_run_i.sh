#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log
tail -8 gpurun_out/r02i_pytest.log | cut -c1-300
python scripts/quick_time.py 2>&1 | tail -2
python scripts/quick_train.py 2>&1 | tail -1
