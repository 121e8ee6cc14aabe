#!/bin/bash
# GPU-box script: correctness + timing of the Cholesky variants of the tuning build (tools/build_dev.sh).
# usage: tools/gpu_chol_sweep.sh "<variants>" [pytest-variants]
cd "$(dirname "$0")/.."
PKG=semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200
DEV=$PWD/$PKG/libsbce_dev.so
VARS=${1:-"2 30 31 32 33 34 35"}
PYV=${2:-"30 31"}
for v in $VARS; do
  echo "== SBCE_CHOL=$v"
  SBCE_LIBRARY=$DEV SBCE_CHOL=$v timeout 300 tools/bench_kernels.sh
done
for v in $PYV; do
  echo "== pytest with SBCE_CHOL=$v"
  SBCE_LIBRARY=$DEV SBCE_CHOL=$v timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "mstep or em_batch or golden or north_star" 2>&1 | tail -4
done
