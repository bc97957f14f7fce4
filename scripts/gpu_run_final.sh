#!/bin/bash
# Evidence run (one GPU): GPU tests, then the headline bench line (config [1] legs only; the driver's round-end run carries the other configs).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -20 gpurun_out/r02f_pytest.log | cut -c1-200
( time timeout 150 python bench.py --configs 1 --no-cpu-baseline > gpurun_out/r02_bench_n1_headline.json 2> gpurun_out/r02_bench_n1_headline.err ) 2>&1 | grep real
cut -c1-2500 gpurun_out/r02_bench_n1_headline.json
