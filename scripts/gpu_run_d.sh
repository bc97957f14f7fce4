#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
tail -25 gpurun_out/r02d_pytest.log
( time python bench.py --steps 50 --warmup 5 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err ) 2>&1 | tail -3
tail -5 gpurun_out/r02d_bench_n1.err; cat gpurun_out/r02d_bench_n1.json | head -c 6000
