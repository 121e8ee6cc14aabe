#!/bin/bash
# GPU-box script for an N-GPU box: the torchrun bench line (NCCL reduce in the timed sweep legs), the test that
# needs two devices, and the INT8 library-GEMM rate quoted in DESIGN.md section 9.
# usage: tools/gpu_multi.sh <tag> <ngpus>
cd "$(dirname "$0")/.."
T=${1:-r02}; N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -c 400 gpurun_out/${T}_bench_${N}gpu.err
python -m pytest tests/test_gpu_configs.py -q -k "non_current or partitioned or chunked" 2>&1 | tail -3
python tools/int8_peak.py > gpurun_out/${T}_int8_peak.json 2>&1; cat gpurun_out/${T}_int8_peak.json
tools/bench_kernels.sh
