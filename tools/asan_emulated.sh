#!/bin/bash
# The whole library (wolken_b200.cu: ABI, host orchestration, every kernel) built for the host over the SIMT emulator
# with -fsanitize=address, and the GPU parity tests run against it: a memcheck of host and device code without a GPU.
# Takes about four minutes.  Usage: tools/asan_emulated.sh [pytest -k expression]
# (For UBSan: the same g++ line with -fsanitize=undefined -fno-sanitize=vptr, LD_PRELOAD=$(gcc -print-file-name=libubsan.so),
#  and pytest -s so that the reports are not captured.)
set -e
cd "$(dirname "$0")/.."
make -s -C tests/simt libwolken_b200_emulated.so        # (re)generates tests/simt/gen
( cd tests/simt && g++ -std=gnu++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared \
    -pthread -Ishim -I. -Igen -Wno-unused-function -o /tmp/libwolken_emul_asan.so gen/wolken_b200.cu.cpp )
K=${1:-"pipeline_matches_oracle and (5000 or 20000) or ragged_and_tiny or identical_locations or encode_same_layout or street_30k_tile3 or return_number_zero or patch_records or error_behaviour or injected_tile or store_queries or formats_and_multifile or different_offsets"}
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
  WB_LIB=/tmp/libwolken_emul_asan.so python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$K"
