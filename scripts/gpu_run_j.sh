#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG"; exit 1; }
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest.log
tail -8 gpurun_out/r02j_pytest.log | cut -c1-300
timeout 120 python scripts/quick_time.py 2>&1 | tail -2
timeout 120 python scripts/quick_train.py 2>&1 | tail -1
