#!/bin/bash
# One bounded GPU pass: the tests that cover the latest changes first, then the default bench line,
# then the rest of the GPU suite.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
( time timeout 170 python -m pytest tests/test_abi.py tests/test_host_cli.py -q -x -k "abi or embuffer or reference_mode or exports" ) > gpurun_out/t1.log 2>&1
tail -3 gpurun_out/t1.log
( time timeout 200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
( time timeout 400 python -m pytest tests -m gpu -q -x ) > gpurun_out/t2.log 2>&1
tail -5 gpurun_out/t2.log
