#!/bin/bash
# BASELINE.json configs 3, 4, 5 through bench.py on one B200 (same JSON schema as the headline)
mkdir -p gpurun_out
for c in 4 3; do
  timeout 900 python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_config$c.log 2>&1; echo config $c rc=$?
  tail -1 gpurun_out/bench_config$c.log > gpurun_out/bench_config$c.json
done
timeout 1500 python bench.py --config 5 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_config5.log 2>&1; echo config 5 rc=$?
tail -1 gpurun_out/bench_config5.log > gpurun_out/bench_config5.json
for c in 4 3 5; do python - $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_config{sys.argv[1]}.json"))
    print(sys.argv[1], d["value"], d["e2e"]["value"], d["config"]["pbs_per_image"], d["check"]["max_abs_deviation_from_clear"], d["check"]["clear_output_span"], d["kernel_breakdown_s_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
