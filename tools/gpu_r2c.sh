#!/bin/bash
# Round-2 GPU pass on ONE B200: the GPU suite, the default bench line, and the sharded pipeline at full size with
# 4 ranks as threads on one device (stage times + labels against one context holding the whole cloud).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2c_tests.log 2>&1
tail -5 gpurun_out/r2c_tests.log
( time timeout 300 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 1500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
( time timeout 600 python tools/shard_check.py --world 4 --points 125000000 --compare-single --steps 2 ) > gpurun_out/r2c_shard4.json 2> gpurun_out/r2c_shard4.err
tail -c 3000 gpurun_out/r2c_shard4.json; tail -3 gpurun_out/r2c_shard4.err
