#!/bin/bash
# Four GPUs, bounded: the 4-GPU bench line as the driver launches it.
mkdir -p gpurun_out
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus 4 --steps 3 --warmup 3 ) > gpurun_out/r2m_bench4.json 2> gpurun_out/r2m_bench4.err
tail -c 1200 gpurun_out/r2m_bench4.json; grep -v "^ *File\|frame #\|^\*\|OMP_NUM" gpurun_out/r2m_bench4.err | tail -5
