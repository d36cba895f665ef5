#!/bin/bash
# One GPU: the scenes again after the adaptive scan bundle, and the census kernel's time.
mkdir -p gpurun_out
for sp in 4:50000000 1:10000000 5:50000000; do
  timeout 200 python bench.py --scene ${sp%%:*} --points ${sp##*:} --steps 3 --warmup 3 --no-cpu --no-scaling-base \
      > gpurun_out/r2t_scene${sp%%:*}.json 2> gpurun_out/r2t_scene${sp%%:*}.err
  python - ${sp%%:*} <<'PY'
import json, sys
l=json.loads(open('gpurun_out/r2t_scene%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(l['config']['workload'][:30], '%.3e points/s' % l['value'], '%.1f ms' % l['ms_per_step'], 'e2e %.1f ms' % l['e2e']['ms_per_step'], l['phases_ms'])
PY
done
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print('C2 resident %.1f' % l['ms_per_step'], 'e2e pipelined %.1f' % l['e2e']['ms_per_step'], 'serial %.1f' % l['e2e']['serial']['ms_per_step'], l['roofline']['traffic'], l['phases_ms'])
print('strip8', l['weak_scaling_base']['ms_per_step'], l['weak_scaling_base']['phases_ms'])
PY
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -x -k "census or full_parity or compiled_reference" ) 2>&1 | tail -2
