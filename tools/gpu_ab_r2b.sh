#!/bin/bash
# Round-2, second A/B: classify variants on BOTH regimes — the bench tile (C2, hyperboloids of tens of metres) and one
# strip of the 8-GPU scene (C3, hyperboloids of a few metres).  Build first:
#   tools/build_variants.sh "r1:-DWB_CL_XWANTS=0 -DWB_CL_REFILTER=0 -DWB_CL_COMPACT2=0 -DWB_CL_NORELVOTE=0 -DWB_CL_F4=0" ...
mkdir -p gpurun_out
for v in ${VARIANTS:-t txl1 tall t30 t16 xl1}; do
  WB_LIB=$PWD/build/variants/lib_$v.so timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu --no-scaling-base \
      > gpurun_out/r2b_c2_$v.json 2> gpurun_out/r2b_c2_$v.err
  WB_LIB=$PWD/build/variants/lib_$v.so timeout 240 python bench.py --strip 8:0 --steps 2 \
      > gpurun_out/r2b_s8_$v.json 2> gpurun_out/r2b_s8_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    a = json.loads(open("gpurun_out/r2b_c2_%s.json" % v).read().strip().splitlines()[-1])
    b = json.loads(open("gpurun_out/r2b_s8_%s.json" % v).read().strip().splitlines()[-1])
    print(v, "C2 classify %.1f ms" % a["phases_ms"]["classify_kernel"], a["labels"]["ground"],
          "| strip8 classify %.1f ms step %.1f" % (b["classify_kernel_ms"], b["ms_per_step"]), b["labels"]["ground"],
          "nodes %.1f" % b["classify_work"]["nodes"])
except Exception as e:
    print(v, "failed", e)
PY
done
