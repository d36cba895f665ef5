#!/bin/bash
# One B200: the whole GPU suite (decode now stages by cp.async.bulk), the bench line, then a source-level profile of
# classify<1> on the bench tile with the instruction-count sections only (CSV exported on the box).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2k_tests.log 2>&1
grep -E "passed|failed|error" gpurun_out/r2k_tests.log | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print('C2', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'serial', l['e2e']['serial']['ms_per_step'], l['phases_ms'])
b=l.get('weak_scaling_base',{})
print('strip8', b.get('ms_per_step'), b.get('phases_ms'))
print('decode', l['roofline_decode'])
PY
timeout 600 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --section MemoryWorkloadAnalysis \
    --clock-control none --import-source on -k regex:wb_classify_kernel -c 1 \
    -f -o /tmp/r2k_classify python bench.py --steps 1 --warmup 0 --no-cpu --no-scaling-base > gpurun_out/r2k_classify_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r2k_classify.ncu-rep --page raw --csv > gpurun_out/r2k_classify_raw.csv 2>/dev/null
ncu -i /tmp/r2k_classify.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2k_classify_source.csv 2>/dev/null
ls -la gpurun_out/r2k_*; du -sh gpurun_out
