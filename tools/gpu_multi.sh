#!/bin/bash
# Multi-GPU box session: the two multi-device tests, then bench.py under torchrun at N = $1 (default 2).
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu or in_process_multi_device or cell_batch_multi_device" 2>&1 | tail -30) > gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${2:-5} --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -c 6000 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
