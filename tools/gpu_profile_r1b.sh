#!/bin/bash
# Launch list of the default bench command with the current kernels (shares, not absolutes), then one
# short bench line per remaining BASELINE scene (C1 street 10M, C4 terrestrial 50M, C5 steep urban 50M).
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/launches_r1b.log 2>&1
tail -c 300 gpurun_out/launches_r1b.log
for cfg in "1 10000000" "4 50000000" "5 50000000"; do
  set -- $cfg
  timeout 70 python bench.py --scene $1 --points $2 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_scene$1.json 2> gpurun_out/bench_scene$1.err
  echo "scene $1 rc=$?"; tail -c 200 gpurun_out/bench_scene$1.json
done
