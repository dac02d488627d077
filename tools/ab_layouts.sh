#!/bin/bash
# Accumulator layouts against each other on one B200 (2 timed images each): default, per-channel offsets only, the Concrete-like
# tensor-wide layout, and the opt-in fused residual lookups.  Outputs: gpurun_out/ab_<variant>.json
mkdir -p gpurun_out
for v in ${@:-default nowidths tensorwide fused}; do
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
    print(sys.argv[1], d["value"], d["e2e"]["value"], d["config"]["pbs_per_image"], d["check"]["max_abs_deviation_from_clear"], d["check"]["clear_output_span"], d["config"]["accumulator_layout"][:40])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
