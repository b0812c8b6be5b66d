#!/bin/bash
# One GPU-box session of round 2: parity tests, bench line, ncu launch lists at the sizes of interest.
# Usage (under gpurun): bash tools/gpu_round2.sh [tests|notests] [sizes for the launch lists ...]
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
if [ "${1:-tests}" = "tests" ]; then
  (timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
shift
timeout 600 python tools/gpu_quick.py 4096 65536 131072 1048576 > gpurun_out/quick.log 2>&1; tail -5 gpurun_out/quick.log | cut -c1-600
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 5000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
for n in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$n.csv python tools/gpu_quick.py $n > gpurun_out/ncu_$n.log 2>&1
done
