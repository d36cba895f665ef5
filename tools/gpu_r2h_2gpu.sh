#!/bin/bash
# Two GPUs: wb_shard_run over real NCCL against the oracle (processes and threads), wolkencli --gpus 2, then the
# 2-GPU bench line as the driver launches it.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_smi.txt
( time timeout 600 python -m pytest tests/test_multigpu_nccl.py tests/test_host_cli.py -m gpu -q -x -k "nccl or cli_gpus" ) > gpurun_out/r2h_nccl_tests.log 2>&1
grep -E "passed|failed|error|skipped" gpurun_out/r2h_nccl_tests.log | tail -3
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2h_bench2.json 2> gpurun_out/r2h_bench2.err
tail -c 3500 gpurun_out/r2h_bench2.json; tail -5 gpurun_out/r2h_bench2.err
