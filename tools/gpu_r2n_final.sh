#!/bin/bash
# Final single-GPU pass of round 2: GPU suite, compute-sanitizer memcheck of a small run through every kernel, the bench
# line, and ncu --set full of (a) classify<1> on the bench tile, (b) the tile-phase and ordering kernels on the C3
# strip, (c) the encode kernel through wolkencli.  Reports stay on the box; raw metrics come back as CSV.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x -s ) > gpurun_out/r2n_tests.log 2>&1
grep -E "passed|failed|error|sampled labels" gpurun_out/r2n_tests.log | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py ) > gpurun_out/r2n_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|sharded|census" gpurun_out/r2n_memcheck.log | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print('C2', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'serial', l['e2e']['serial']['ms_per_step'], l['roofline'])
PY
for sp in 1:10000000 4:50000000 5:50000000; do
  timeout 200 python bench.py --scene ${sp%%:*} --points ${sp##*:} --steps 3 --warmup 3 --no-cpu --no-scaling-base \
      > gpurun_out/r2n_scene${sp%%:*}.json 2> gpurun_out/r2n_scene${sp%%:*}.err
  python - ${sp%%:*} <<'PY'
import json, sys
l=json.loads(open('gpurun_out/r2n_scene%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(l['config']['workload'][:40], '%.3e points/s' % l['value'], '%.1f ms' % l['ms_per_step'], 'e2e %.1f ms' % l['e2e']['ms_per_step'], l['phases_ms'])
PY
done
timeout 600 ncu --set full --clock-control none -k regex:wb_classify_kernel -c 1 -f -o /tmp/r2n_classify \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-scaling-base > gpurun_out/r2n_classify_ncu.log 2>&1
echo "ncu classify rc=$?"
ncu -i /tmp/r2n_classify.ncu-rep --page raw --csv > gpurun_out/r2n_classify_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none \
    -k regex:"^wb_(scan_kernel|member_count|member_fill|segment|postscan|hilbert|classify_gather|classify_scatter|chunk_bounds|tile_extent|tile_grid|max_hyp|scatter_labels|leaf_emit|pending)" \
    -c 20 -f -o /tmp/r2n_tiles python bench.py --strip 8:0 --steps 1 > gpurun_out/r2n_tiles_ncu.log 2>&1
echo "ncu tiles rc=$?"
ncu -i /tmp/r2n_tiles.ncu-rep --page raw --csv > gpurun_out/r2n_tiles_raw.csv 2>/dev/null
python - <<'PY'
import sys
sys.path.insert(0, '.')
from wolkenbase_b200 import synth
synth.generate(1, 10_000_000, seed=1).write('/tmp/c1.las')
PY
timeout 300 ncu --set full --clock-control none -k regex:"wb_encode_kernel|wb_leaf_class_counts|wb_census" -c 4 -f -o /tmp/r2n_encode \
    wolkenbase_b200/host/wolkencli -o /tmp/c1_out --dump /tmp/c1_dump /tmp/c1.las > gpurun_out/r2n_encode_ncu.log 2>&1
echo "ncu encode rc=$?"
ncu -i /tmp/r2n_encode.ncu-rep --page raw --csv > gpurun_out/r2n_encode_raw.csv 2>/dev/null
ls -la gpurun_out/r2n_*; du -sh gpurun_out
