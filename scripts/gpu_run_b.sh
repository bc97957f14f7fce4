#!/bin/bash
# ncu --set full of the two blend kernels (one launch each), source import on
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:blend_ --launch-skip 6 --launch-count 2 \
    -o gpurun_out/r02b_blend -f python scripts/quick_time.py > gpurun_out/r02b_ncu.log 2>&1
tail -3 gpurun_out/r02b_ncu.log
ls -la gpurun_out/r02b_blend.ncu-rep
