#!/bin/bash
# Round-2 opener (run under gpurun after tools/build_variants.sh base: xwants:-DWB_CL_XWANTS=1 refilter:-DWB_CL_REFILTER=1
# compact2:-DWB_CL_COMPACT2=1 "xr:-DWB_CL_XWANTS=1 -DWB_CL_REFILTER=1" "all:-DWB_CL_XWANTS=1 -DWB_CL_REFILTER=1 -DWB_CL_COMPACT2=1"):
# the emulator-validated classify candidates on the bench workload, then the whole GPU suite on the combination.
mkdir -p gpurun_out
for v in base xwants refilter compact2 xr all; do
  WB_LIB=$PWD/build/variants/lib_$v.so timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu --no-scaling-base \
      > gpurun_out/r2_ab_$v.json 2> gpurun_out/r2_ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open("gpurun_out/r2_ab_%s.json" % v).read().strip().splitlines()[0])
print(v, round(d["ms_per_step"], 1), d["phases_ms"]["classify_kernel"], d["labels"], d["classify_work"])
PY
done
( time WB_LIB=$PWD/build/variants/lib_all.so timeout 400 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_ab_tests_all.log 2>&1
tail -4 gpurun_out/r2_ab_tests_all.log
