#!/bin/bash
# GPU-box script: the round's evidence run.  usage: tools/gpu_final.sh <tag>   (e.g. r02i)
cd "$(dirname "$0")/.."
T=${1:-r02}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_reference_arm.err
for c in 1 3 4 41 5; do
  python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; tail -c 200 gpurun_out/${T}_bench_c$c.err
done
# launch list of the default command (only after it exited 0 without ncu), then one full capture of the hot kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k "regex:k_chol_solve|k_gram_tma4|k_enum|k_heff_qr_mma" --launch-skip 24 -c 4 \
    -o gpurun_out/${T}_full -f python bench.py --kernels-only --steps 1 --warmup 3 > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/${T}_*
