#!/bin/bash
# Round-2 GPU pass after the Hilbert classify order: new tests first, the bench line, then the whole GPU suite.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_parity.py tests/test_host_cli.py -m gpu -q -x -s -k "config_scale or census or cli_gpus or full_parity" ) > gpurun_out/r2d_new_tests.log 2>&1
tail -12 gpurun_out/r2d_new_tests.log
( time timeout 300 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 2500 gpurun_out/r2d_bench.json | head -c 1800; echo; tail -3 gpurun_out/r2d_bench.err
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_tests.log 2>&1
tail -5 gpurun_out/r2d_tests.log
