#!/bin/bash
# ncu --set full on the two PBS kernels at the shipped sets (batch 12288); plain run first (must exit 0)
mkdir -p gpurun_out
python tools/microbench.py --batches 12288 --no-ks --iters 1 --warmup 1 > gpurun_out/micro_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pbs_kernel' -s 2 -c 2 -f -o gpurun_out/${1:-pbs_r02} \
    python tools/microbench.py --batches 12288 --no-ks --iters 1 --warmup 1 > gpurun_out/micro_ncu.log 2>&1
tail -3 gpurun_out/micro_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
