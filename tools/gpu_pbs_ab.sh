#!/bin/bash
# PBS kernel A/B on one B200: parity of the kernel tests, then the microbench at the shipped sets for each variant.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_circuit_gpu.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; tail -5 gpurun_out/pytest_kernels.log
TFX_PBS_VERBOSE=1 timeout 200 python tools/microbench.py --batches 12288 --no-ks --json gpurun_out/micro_v8.json 2>&1 | grep -v "^$" | cut -c1-260 | sort -u | tail -8
TFX_PBS_V7=1 timeout 200 python tools/microbench.py --batches 12288 --no-ks --json gpurun_out/micro_v7.json 2>&1 | cut -c1-260 | tail -3
