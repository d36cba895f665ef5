#!/bin/bash
# Round-2 profiling pass (one B200, under gpurun).  Outputs land in gpurun_out/.
#   1. strips of the multi-GPU scenes on one GPU (bench.py --strip W:R): where the 4-GPU run of round 1 went
#   2. ncu --set full of the shipped classify kernels on the bench workload, source page exported on the box
#   3. launch list (gpu__time_duration) of one strip step: the non-classify kernels of the sharded scene
mkdir -p gpurun_out
for s in ${STRIPS:-8:0 4:1 4:0 2:0}; do
  timeout 300 python bench.py --strip $s --steps 2 > gpurun_out/r2_strip_${s/:/_}.json 2> gpurun_out/r2_strip_${s/:/_}.err
  echo "strip $s rc=$?"; tail -c 900 gpurun_out/r2_strip_${s/:/_}.json; echo
done
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wb_classify_kernel -c 2 \
    -o gpurun_out/r2_classify_full python bench.py --steps 1 --warmup 0 --no-cpu --no-scaling-base > gpurun_out/r2_classify_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r2_classify_full.ncu-rep --page raw --csv > gpurun_out/r2_classify_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_classify_full.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2_classify_source.csv 2>/dev/null
ls -la gpurun_out/r2_classify_*
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_strip8.csv python bench.py --strip 8:0 --steps 1 > gpurun_out/r2_launches_strip8.log 2>&1
echo "launch list rc=$?"
fi
