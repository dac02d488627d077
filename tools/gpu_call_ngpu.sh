#!/bin/bash
# N-GPU evidence on one box: headline bench, config 4, the kernel sweep, and the reference arm — each under torch.distributed.run
N=${1:-8}; shift
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N; run 29531 bench.py --gpus $N --steps 3 --warmup 1" > gpurun_out/bench_${N}gpu.log 2>&1; echo bench rc=$?
tail -1 gpurun_out/bench_${N}gpu.log > gpurun_out/bench_${N}gpu.json
timeout 600 bash -c "$(declare -f run); N=$N; run 29532 bench.py --gpus $N --config 4 --steps 2 --warmup 1" > gpurun_out/bench_config4_${N}gpu.log 2>&1; echo config4 rc=$?
tail -1 gpurun_out/bench_config4_${N}gpu.log > gpurun_out/bench_config4_${N}gpu.json
timeout 400 bash tools/gpu_sweep.sh $N
( time timeout 300 bash -c "$(declare -f run); N=$N; run 29533 bench.py --impl reference --gpus $N --steps 3 --warmup 1" ) > gpurun_out/ref_${N}gpu.log 2>&1; echo ref rc=$?
if [ "$1" = "config5" ]; then
  timeout 900 bash -c "$(declare -f run); N=$N; run 29534 bench.py --gpus $N --config 5 --steps 1 --warmup 0" > gpurun_out/bench_config5_${N}gpu.log 2>&1; echo config5 rc=$?
  tail -1 gpurun_out/bench_config5_${N}gpu.log > gpurun_out/bench_config5_${N}gpu.json
fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in (f"bench_{n}gpu", f"bench_config4_{n}gpu", f"bench_config5_{n}gpu"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["output_sha"], d["check"]["max_abs_deviation_from_clear"], {k: round(v, 3) for k, v in d["kernel_breakdown_s_per_step"].items()})
    except Exception as e:
        print(f, "missing", e)
PY
tail -3 gpurun_out/ref_${N}gpu.log | cut -c1-200
