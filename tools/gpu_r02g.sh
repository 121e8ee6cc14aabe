#!/bin/bash
cd "$(dirname "$0")/.."
PKG=semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200
DEV=$PWD/$PKG/libsbce_dev.so
for v in 52 56 57 58; do echo "== SBCE_CHOL=$v"; SBCE_LIBRARY=$DEV SBCE_CHOL=$v timeout 300 tools/bench_kernels.sh; done
echo "== SBCE_CHOL=52 SBCE_ENUM_MINB=3"; SBCE_LIBRARY=$DEV SBCE_CHOL=52 SBCE_ENUM_MINB=3 timeout 300 tools/bench_kernels.sh
echo "== production library"; timeout 300 tools/bench_kernels.sh
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "mstep or em_batch or golden or estep" 2>&1 | tail -3
