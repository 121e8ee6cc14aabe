#!/bin/bash
# GPU-box script: one `ncu --set full` capture (source-level) of selected kernels of the default bench workload.
# usage: tools/gpu_ncu_capture.sh <out-name> <kernel-regex> [launch-skip] [count] [env assignments...]
cd "$(dirname "$0")/.."
OUT=$1; REGEX=$2; SKIP=${3:-12}; CNT=${4:-1}; shift 4 || true
mkdir -p gpurun_out
env "$@" ncu --set full --import-source on --clock-control none -k "regex:$REGEX" --launch-skip "$SKIP" -c "$CNT" \
    -o "gpurun_out/$OUT" -f python bench.py --kernels-only --steps 1 --warmup 3 > "gpurun_out/$OUT.log" 2>&1
tail -3 "gpurun_out/$OUT.log"
ls -la gpurun_out/$OUT.ncu-rep
