#!/bin/bash
# A/B of the classify kernel's single-precision shortcuts (build/variants/lib_<FSECTOR><FREACH><FSPAN>.so, built by
# hand with -DWB_CL_...=0/1) on the default bench workload, then the whole GPU suite on the default library.
mkdir -p gpurun_out
WB_LIB=$PWD/build/variants/lib_000.so timeout 80 python bench.py --steps 2 --warmup 1 --no-cpu --no-scaling-base > gpurun_out/ab_000.json 2> gpurun_out/ab_000.err
python - <<'PY'
import json
for v in ("000",):
    d=json.loads(open("gpurun_out/ab_%s.json"%v).read().strip().splitlines()[0])
    print(v, round(d["ms_per_step"],1), d["phases_ms"]["classify_kernel"], d["labels"], d["classify_work"]["pair_tests_per_point"])
PY
timeout 90 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ab_111.json 2> gpurun_out/ab_111.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ab_111.json").read().strip().splitlines()[0])
print("111", round(d["ms_per_step"],1), d["phases_ms"]["classify_kernel"], d["labels"], d["classify_work"]["pair_tests_per_point"], d.get("weak_scaling_base"))
PY
( time timeout 300 python -m pytest tests -m gpu -q -x ) > gpurun_out/ab_tests.log 2>&1
tail -4 gpurun_out/ab_tests.log
WB_LIB=$PWD/build/variants/lib_110.so timeout 60 python bench.py --steps 2 --warmup 1 --no-cpu --no-scaling-base > gpurun_out/ab_110.json 2> gpurun_out/ab_110.err
tail -c 100 gpurun_out/ab_110.json
