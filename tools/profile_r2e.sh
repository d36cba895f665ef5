#!/bin/bash
# ncu --set full of every kernel of one step EXCEPT classify, on rank 0's strip of the 8-GPU scene (C3: small tiles,
# where scan, membership pairs and the sorts weigh most), plus the launch list of the same command.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2e_tests.log 2>&1
tail -4 gpurun_out/r2e_tests.log
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"^wb_(scan|member|segment|postscan|sort|split|keygen|decode|hilbert|classify_gather|classify_scatter|chunk_bounds|node_bounds|leaf|coords|gather|scatter_labels|tile)" \
    -c 60 -f -o gpurun_out/r2e_strip8_full python bench.py --strip 8:0 --steps 1 > gpurun_out/r2e_strip8_full.log 2>&1
echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2e_launches_strip8.csv python bench.py --strip 8:0 --steps 1 > gpurun_out/r2e_launches_strip8.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/r2e_*
