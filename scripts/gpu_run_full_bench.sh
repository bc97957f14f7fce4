#!/bin/bash
# The default bench line (every config leg + cpu_baseline), as the driver runs it at round end.
mkdir -p gpurun_out
( time timeout 140 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
grep "\[bench\]" gpurun_out/r02_bench_n1.err | tail -4
cut -c1-200 gpurun_out/r02_bench_n1.json
