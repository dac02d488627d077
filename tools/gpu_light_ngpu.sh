N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_${N}gpu.log > gpurun_out/bench_${N}gpu.json
bash tools/gpu_sweep.sh $N > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print($N, d['value'], d['e2e']['value'], d['output_sha'], d['kernel_breakdown_s_per_step'].get('gather'))"
