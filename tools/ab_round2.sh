#!/bin/bash
# First GPU call of the next round (about 5 GPU-minutes on one B200): validates the opt-in paths that round 1 could only check on
# the CPU and times the accumulator layouts against each other.  Outputs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/ab_round2.sh'
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
# default layout (per-channel offsets + widths): never timed in round 1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default.log 2>&1; tail -1 gpurun_out/ab_default.log > gpurun_out/ab_default.json
# previous layout (per-channel offsets only)
TFX_PER_CHANNEL_WIDTHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ab_nowidths.log 2>&1; tail -1 gpurun_out/ab_nowidths.log > gpurun_out/ab_nowidths.json
# opt-in fused residual lookups
TFX_FUSE_RESIDUAL=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ab_fused.log 2>&1; tail -1 gpurun_out/ab_fused.log > gpurun_out/ab_fused.json
for f in default nowidths fused; do python - "$f" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
print(sys.argv[1], d["value"], d["config"]["pbs_per_image"], d["check"], d["kernel_breakdown_s_per_step"])
PY
done
