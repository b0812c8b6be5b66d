#!/bin/bash
# One GPU-box session: parity tests, quick stage timings, bench line, ncu launch list + one full capture.
# Usage (under gpurun): bash tools/gpu_round.sh [tests|notests] [ncu|noncu] [kernel-regex]
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
if [ "${1:-tests}" = "tests" ]; then
  (timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
timeout 600 python tools/gpu_quick.py 4096 65536 1048576 > gpurun_out/quick.log 2>&1; tail -4 gpurun_out/quick.log | cut -c1-700
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 4500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.log
if [ "${2:-ncu}" = "ncu" ]; then
  K="${3:-k_subgroup_chain2}"
  timeout 300 python tools/gpu_quick.py 65536 > gpurun_out/plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python tools/gpu_quick.py 65536 > gpurun_out/ncu1.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_$K python tools/gpu_quick.py 65536 > gpurun_out/ncu2.log 2>&1
  tail -3 gpurun_out/ncu2.log
fi
