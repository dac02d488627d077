#!/bin/bash
# One profiling round on a B200 box: plain runs first (must exit 0), then ncu.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/pytest_kernels.log 2>&1 || { tail -20 gpurun_out/pytest_kernels.log; exit 1; }
python tools/microbench.py --batches 1184,12288 --json gpurun_out/micro.json > gpurun_out/micro.log 2>&1 || exit 1
tail -5 gpurun_out/micro.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 || { tail -5 gpurun_out/bench_plain.log; exit 1; }
tail -1 gpurun_out/bench_plain.log | cut -c1-300
python tools/microbench.py --batches 12288 --iters 1 --warmup 1 > gpurun_out/micro_plain2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pbs_kernel|keyswitch_kernel' -s 4 -c 4 -o gpurun_out/kernels_full \
    python tools/microbench.py --batches 12288 --iters 1 --warmup 1 > gpurun_out/micro_ncu.log 2>&1
tail -2 gpurun_out/micro_ncu.log
ls -la gpurun_out
