#!/bin/bash
# Round-1 profiling pass (run under gpurun on one B200).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
P=${1:-100000000}
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --points $P --steps 1 --warmup 0 --no-cpu > gpurun_out/launches_r1.log 2>&1
# the top kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:wb_classify -c 1 \
    -o gpurun_out/classify_r1_full python bench.py --points $P --steps 1 --warmup 0 --no-cpu > gpurun_out/classify_r1_full.log 2>&1
# the HBM-bound kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wb_sort_downsweep -s 3 -c 1 \
    -o gpurun_out/downsweep_r1_full python bench.py --points $P --steps 1 --warmup 0 --no-cpu > gpurun_out/downsweep_r1_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wb_decode -c 1 \
    -o gpurun_out/decode_r1_full python bench.py --points $P --steps 1 --warmup 0 --no-cpu > gpurun_out/decode_r1_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_r1.csv
