#!/bin/bash
# BASELINE.json configs[1]: batched PBS + keyswitch sweep, batch 1..65536, at the shipped parameter sets (1 GPU here; N GPUs under torchrun)
mkdir -p gpurun_out
N=${1:-1}
B=1,2,4,8,16,32,64,128,256,512,1024,2048,4096,8192,16384,32768,65536
if [ "$N" = "1" ]; then
  python tools/microbench.py --batches $B --iters 2 --warmup 1 --json gpurun_out/sweep_${N}gpu.json > gpurun_out/sweep_${N}gpu.log 2>&1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/microbench.py --batches $B --iters 2 --warmup 1 --json gpurun_out/sweep_${N}gpu.json > gpurun_out/sweep_${N}gpu.log 2>&1
fi
echo rc=$?; tail -4 gpurun_out/sweep_${N}gpu.log | cut -c1-300
