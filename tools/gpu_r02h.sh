#!/bin/bash
cd "$(dirname "$0")/.."
PKG=semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200
DEV=$PWD/$PKG/libsbce_dev.so
echo "== production library"; timeout 300 tools/bench_kernels.sh
echo "== SBCE_ENUM_MINB=3"; SBCE_LIBRARY=$DEV SBCE_ENUM_MINB=3 timeout 300 tools/bench_kernels.sh
(time timeout 1500 python -m pytest tests -m gpu -q) 2>&1 | tail -8
