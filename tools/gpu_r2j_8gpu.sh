#!/bin/bash
# Eight GPUs, bounded: (1) where does host->device bandwidth go when 8 ranks copy at once (tools/h2d_probe.py),
# (2) the 8-GPU bench line as the driver launches it.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_smi.txt
lscpu | head -25 > gpurun_out/r2j_lscpu.txt 2>&1
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    tools/h2d_probe.py ) > gpurun_out/r2j_probe.txt 2> gpurun_out/r2j_probe.err
head -c 2500 gpurun_out/r2j_probe.txt; grep -v "^ *File\|frame #\|^\*\|OMP_NUM" gpurun_out/r2j_probe.err | tail -5
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2j_bench8.json 2> gpurun_out/r2j_bench8.err
tail -c 3000 gpurun_out/r2j_bench8.json; grep -v "^ *File\|frame #\|^\*\|OMP_NUM" gpurun_out/r2j_bench8.err | tail -8
