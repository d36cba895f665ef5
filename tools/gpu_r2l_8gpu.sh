#!/bin/bash
# Eight GPUs, bounded: the 8-GPU bench line as the driver launches it (after the priority load stream and the
# Hilbert-ordered second store), then wolkencli --gpus 8 on 8 strip files (file to file).
mkdir -p gpurun_out
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2l_bench8.json 2> gpurun_out/r2l_bench8.err
tail -c 1500 gpurun_out/r2l_bench8.json; grep -v "^ *File\|frame #\|^\*\|OMP_NUM" gpurun_out/r2l_bench8.err | tail -5
( time timeout 300 python tools/cli_gpus_check.py 8 40000000 ) > gpurun_out/r2l_cli8.txt 2>&1
tail -25 gpurun_out/r2l_cli8.txt
