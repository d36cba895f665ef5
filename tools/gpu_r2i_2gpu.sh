#!/bin/bash
# Two GPUs: the 2-GPU bench line as the driver launches it (bounded), then the N=2 reference arm's shape check.
mkdir -p gpurun_out
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err
tail -c 4000 gpurun_out/r2i_bench2.json; grep -v "^ *File\|^    \|frame #" gpurun_out/r2i_bench2.err | tail -8
