#!/bin/bash
# 2-GPU call: N-GPU == 1-GPU word parity, bench at N=2 (torchrun), reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -rs > gpurun_out/pytest_multigpu.log 2>&1; tail -4 gpurun_out/pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu.log 2>&1
tail -1 gpurun_out/bench_2gpu.log > gpurun_out/bench_2gpu.json; tail -1 gpurun_out/bench_2gpu.log | cut -c1-400
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/ref_2gpu.log 2>&1
tail -5 gpurun_out/ref_2gpu.log | cut -c1-700
