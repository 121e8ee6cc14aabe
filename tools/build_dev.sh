#!/bin/bash
# Tuning build: libsbce_dev.so = the same sources compiled with -DSBCE_DEV, which enables run-time kernel-variant
# knobs read from the environment (common.cuh: dev_knob).  Never shipped or loaded by default; select it with
#   SBCE_LIBRARY=$PWD/<package>/libsbce_dev.so SBCE_CHOL=31 python bench.py --kernels-only
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200"
OUT="$PKG/csrc/dev_build"
mkdir -p "$OUT"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DSBCE_DEV"
pids=()
for f in abi estep mstep chol metrics pm gen; do
  nvcc $FLAGS -c "$PKG/csrc/$f.cu" -o "$OUT/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$PKG/libsbce_dev.so" "$OUT"/*.o -lcudart
echo "built $PKG/libsbce_dev.so"
