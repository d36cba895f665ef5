#!/bin/bash
# Eight GPUs: does pacing the chunk copies (at most two queued) help the pipelined e2e?  Same bench line twice.
mkdir -p gpurun_out
for paced in 1 0; do
  WB_H2D_PACED=$paced timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$paced \
      bench.py --gpus 8 --steps 3 --warmup 3 --no-scaling-base > gpurun_out/r2p_bench8_paced$paced.json 2> gpurun_out/r2p_bench8_paced$paced.err
  python - $paced <<'PY'
import json, sys
l=json.loads(open('gpurun_out/r2p_bench8_paced%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('paced', sys.argv[1], 'resident %.1f ms' % l['ms_per_step'], 'e2e pipelined %.1f ms' % l['e2e']['ms_per_step'], 'serial %.1f' % l['e2e']['serial']['ms_per_step'], l['e2e']['h2d_decode_ms'], l['parity']['ok'])
PY
done
