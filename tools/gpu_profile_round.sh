#!/bin/bash
# One profiling round on a B200 box: plain runs first (must exit 0), then ncu.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 || { tail -20 gpurun_out/pytest_gpu.log; exit 1; }
tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 || { tail -5 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log
python tools/microbench.py --batches 1,16,148,1184,12288,65536 --json gpurun_out/micro.json > gpurun_out/micro.log 2>&1 || exit 1
tail -3 gpurun_out/micro.log | cut -c1-200
python bench.py > gpurun_out/bench_plain.log 2>&1 || { tail -5 gpurun_out/bench_plain.log; exit 1; }
tail -1 gpurun_out/bench_plain.log > gpurun_out/bench_default.json
cut -c1-200 gpurun_out/bench_default.json
python tools/microbench.py --batches 12288 --iters 1 --warmup 1 > gpurun_out/micro_plain2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pbs_kernel|keyswitch_kernel' -s 1 -c 6 -o gpurun_out/kernels_full \
    python tools/microbench.py --batches 12288 --iters 1 --warmup 1 > gpurun_out/micro_ncu.log 2>&1
tail -2 gpurun_out/micro_ncu.log
ls -la gpurun_out
