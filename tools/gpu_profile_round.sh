#!/bin/bash
# One profiling round on a B200 box: plain runs first (must exit 0), then ncu.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo launches rc=$?; wc -l gpurun_out/bench_launches.csv
python tools/microbench.py --batches 12288 --no-ks --sets tlu --iters 1 --warmup 1 > gpurun_out/micro_plain_tlu.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pbs_kernel' -s 1 -c 1 -f -o gpurun_out/pbs_r02_v8_tlu \
    python tools/microbench.py --batches 12288 --no-ks --sets tlu --iters 1 --warmup 1 > gpurun_out/micro_ncu_tlu.log 2>&1
echo ncu tlu rc=$?
python tools/microbench.py --batches 12288 --iters 1 --warmup 1 > gpurun_out/micro_plain_ks.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'keyswitch_umma' -s 1 -c 1 -f -o gpurun_out/ks_r02_umma \
    python tools/microbench.py --batches 12288 --iters 1 --warmup 1 > gpurun_out/micro_ncu_ks.log 2>&1
echo ncu ks rc=$?
ls -la gpurun_out/*.ncu-rep
