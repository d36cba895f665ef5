#!/bin/bash
# One GPU: does the size of the segment table uploaded by wb_build decide whether the pipelined e2e overlaps?
mkdir -p gpurun_out
for v in seg2048 seg256; do
  WB_LIB=$PWD/build/variants/lib_$v.so timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-scaling-base > gpurun_out/r2q_$v.json 2> gpurun_out/r2q_$v.err
  python - $v <<'PY'
import json, sys
l=json.loads(open('gpurun_out/r2q_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'resident %.1f' % l['ms_per_step'], 'e2e pipelined %.1f' % l['e2e']['ms_per_step'], 'serial %.1f' % l['e2e']['serial']['ms_per_step'])
PY
done
