#!/bin/bash
# GPU-box script: the round's evidence run, in two parts so that each fits a short gpurun call.
#   tools/gpu_final.sh <tag> a   -- GPU tests, default bench, reference arm, the other BASELINE.json configs
#   tools/gpu_final.sh <tag> b   -- ncu launch list of the default command + one full capture of the hot kernels
cd "$(dirname "$0")/.."
T=${1:-r02}
PART=${2:-ab}
mkdir -p gpurun_out
if [[ $PART == *a* ]]; then
  (time python -m pytest tests -m gpu -q) > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
  python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_reference_arm.err
  for c in 1 3 4 41 5; do
    python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; tail -c 200 gpurun_out/${T}_bench_c$c.err
  done
fi
if [[ $PART == *b* ]]; then
  # launch list of the default command (only after it exited 0 without ncu), then one full capture of the hot kernels
  ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_LAUNCHES:-260} --csv --log-file gpurun_out/${T}_launches.csv \
      python bench.py --steps 1 --warmup 3 --kernels-only > gpurun_out/${T}_ncu1.log 2>&1
  ncu --set full --import-source on --clock-control none -k "regex:${NCU_KERNELS:-k_chol_solve|k_gram_tma4|k_enum|k_heff_qr_mma}" --launch-skip ${NCU_SKIP:-24} -c ${NCU_COUNT:-4} \
      -o gpurun_out/${T}_full -f python bench.py --kernels-only --steps 1 --warmup 3 > gpurun_out/${T}_ncu2.log 2>&1
fi
ls -la gpurun_out/${T}_*
