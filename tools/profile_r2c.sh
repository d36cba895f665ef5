#!/bin/bash
# ncu --set full of the shipped classify kernels on the bench workload (C2, 100 M points); source page exported here
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wb_classify_kernel -c 2 \
    -f -o gpurun_out/r2c_classify_full python bench.py --steps 1 --warmup 0 --no-cpu --no-scaling-base > gpurun_out/r2c_classify_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r2c_classify_full*
