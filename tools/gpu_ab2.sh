#!/bin/bash
# Last A/B of the round: the single-precision pair test (build/variants/lib_1111.so = -DWB_CL_FPAIR=1).
mkdir -p gpurun_out
export WB_LIB=$PWD/build/variants/lib_1111.so
timeout 40 python bench.py --steps 2 --warmup 1 --no-cpu --no-scaling-base > gpurun_out/ab_1111.json 2> gpurun_out/ab_1111.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ab_1111.json").read().strip().splitlines()[0])
print("1111", round(d["ms_per_step"],1), d["phases_ms"]["classify_kernel"], d["labels"], d["classify_work"]["pair_tests_per_point"])
PY
( time timeout 45 python -m pytest tests/test_gpu_parity.py -q -x -k "pipeline_matches or compiled_reference_fixture or nondefault or identical_locations or (baseline_scenes and not 5-300000)" ) > gpurun_out/ab2_tests.log 2>&1
tail -3 gpurun_out/ab2_tests.log
( time timeout 40 python -m pytest tests/test_gpu_large.py -q -x -k "2-3000000" ) > gpurun_out/ab2_large.log 2>&1
tail -3 gpurun_out/ab2_large.log
