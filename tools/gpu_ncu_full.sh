#!/bin/bash
# ncu --set full captures of the two hot kernels at the benchmarked size (after a plain run of the same command exited 0).
set -u
mkdir -p gpurun_out
N=${1:-1048576}
timeout 300 python tools/gpu_quick.py $N > gpurun_out/plain_$N.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$N.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decompress_sqrt -s 1 -c 1 -f -o gpurun_out/r2_k1a_n$N python tools/gpu_quick.py $N > gpurun_out/ncu_k1a.log 2>&1
tail -2 gpurun_out/ncu_k1a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm_chunk_pass1_staged -s 3 -c 1 -f -o gpurun_out/r2_k6_n$N python tools/gpu_quick.py $N > gpurun_out/ncu_k6.log 2>&1
tail -2 gpurun_out/ncu_k6.log
ls -la gpurun_out/*.ncu-rep
