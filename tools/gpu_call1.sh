#!/bin/bash
# Round 2, GPU call 1: data-path probes, parity suite, accumulator-layout A/B (profiles/r02_*).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 120 exp/mio_paths > gpurun_out/mio_paths.txt 2>&1; tail -60 gpurun_out/mio_paths.txt
timeout 600 python -m pytest tests -m gpu -q -rxXs > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
for v in default nowidths tensorwide fused; do
  case $v in
    default) E="";;
    nowidths) E="TFX_PER_CHANNEL_WIDTHS=0";;
    tensorwide) E="TFX_PER_CHANNEL_OFFSETS=0";;
    fused) E="TFX_FUSE_RESIDUAL=1";;
  esac
  env $E timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  tail -1 gpurun_out/ab_$v.log > gpurun_out/ab_$v.json
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
    print(sys.argv[1], d["value"], d["config"]["pbs_per_image"], d["check"], d["kernel_breakdown_s_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
