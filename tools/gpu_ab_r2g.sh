#!/bin/bash
# A/B of classify variants after the Hilbert order: the bench tile (C2) and rank 0's strip of the 8-GPU scene (C3).
# Build first with tools/build_variants.sh (build/ travels with the snapshot).
mkdir -p gpurun_out
for v in ${VARIANTS:-pw0 pw1 pw2 pw1t pw1w2 mb24 mb28}; do
  WB_LIB=$PWD/build/variants/lib_$v.so timeout 240 python bench.py --steps 2 --warmup 1 --no-cpu --no-scaling-base \
      > gpurun_out/r2g_c2_$v.json 2> gpurun_out/r2g_c2_$v.err
  WB_LIB=$PWD/build/variants/lib_$v.so timeout 240 python bench.py --strip 8:0 --steps 2 \
      > gpurun_out/r2g_s8_$v.json 2> gpurun_out/r2g_s8_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    a = json.loads(open("gpurun_out/r2g_c2_%s.json" % v).read().strip().splitlines()[-1])
    b = json.loads(open("gpurun_out/r2g_s8_%s.json" % v).read().strip().splitlines()[-1])
    print(v, "C2 classify %.1f ms step %.1f" % (a["phases_ms"]["classify_kernel"], a["ms_per_step"]), a["labels"]["ground"],
          "| strip8 classify %.1f ms step %.1f scan %.1f" % (b["classify_kernel_ms"], b["ms_per_step"], b["phases_ms"]["scan"]),
          b["labels"]["ground"], "nodes %.1f" % b["classify_work"]["nodes"])
except Exception as e:
    print(v, "failed", e)
PY
done
