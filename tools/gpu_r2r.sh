#!/bin/bash
# One GPU: the bench line after the small-message fix (does the pipelined e2e overlap again?) + the sharded tests.
mkdir -p gpurun_out
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-scaling-base > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1])
print('resident %.1f' % l['ms_per_step'], 'e2e pipelined %.1f' % l['e2e']['ms_per_step'], 'serial %.1f' % l['e2e']['serial']['ms_per_step'])
PY
( timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_host_cli.py tests/test_gpu_large.py -m gpu -q -x -k "sharded or cli_gpus or many_files or header_offsets or identical" ) 2>&1 | tail -2
