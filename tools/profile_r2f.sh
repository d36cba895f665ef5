#!/bin/bash
# Full GPU suite, the bench line, rank 0's strip of the 8-GPU scene, then ncu --set full of every kernel of one strip
# step EXCEPT classify (C3: small tiles, where scan, membership pairs and the sorts weigh most).  The report stays on
# the box (it exceeds what gpurun brings back): raw metrics and the scan kernel's source page are exported as CSV.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2f_tests.log 2>&1
grep -E "passed|failed|error" gpurun_out/r2f_tests.log | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print('C2', l['ms_per_step'], l['e2e']['ms_per_step'], l['phases_ms'])
b=l.get('weak_scaling_base',{})
print('strip8', b.get('ms_per_step'), b.get('phases_ms'))
PY
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"^wb_(scan|member|segment|postscan|sort|split|keygen|decode|hilbert|classify_gather|classify_scatter|chunk_bounds|leaf|coords|gather|scatter_labels|tile)" \
    -c 45 -f -o /tmp/r2f_strip8_full python bench.py --strip 8:0 --steps 1 > gpurun_out/r2f_strip8_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/r2f_strip8_full.ncu-rep --page raw --csv > gpurun_out/r2f_strip8_raw.csv 2>/dev/null
ncu -i /tmp/r2f_strip8_full.ncu-rep --page source --csv --print-source cuda,sass -k regex:wb_scan_kernel > gpurun_out/r2f_scan_source.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2f_launches_strip8.csv python bench.py --strip 8:0 --steps 1 > gpurun_out/r2f_launches_strip8.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/r2f_*
du -sh gpurun_out
