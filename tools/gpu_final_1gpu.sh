#!/bin/bash
# Final single-GPU evidence of a round: parity suite, smoke, headline bench (with the CPU baseline), kernel sweep, accumulator-layout
# A/B, configs 4 and 3, then the ncu captures (each after its plain run exited 0).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log > gpurun_out/bench_1gpu.json
bash tools/gpu_sweep.sh 1
bash tools/ab_layouts.sh nowidths tensorwide fused
for c in 4 3; do
  timeout 900 python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_config$c.log 2>&1
  tail -1 gpurun_out/bench_config$c.log > gpurun_out/bench_config$c.json
done
python tools/microbench.py --batches 12288 --no-ks --iters 1 --warmup 1 > gpurun_out/micro_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pbs_kernel' -s 1 -c 1 -f -o gpurun_out/pbs_final_tlu \
    python tools/microbench.py --batches 12288 --no-ks --iters 1 --warmup 1 > gpurun_out/micro_ncu_tlu.log 2>&1
python tools/microbench.py --batches 12288 --no-ks --sets bit --iters 1 --warmup 1 > gpurun_out/micro_plain_bit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pbs_kernel' -s 1 -c 1 -f -o gpurun_out/pbs_final_bit \
    python tools/microbench.py --batches 12288 --no-ks --sets bit --iters 1 --warmup 1 > gpurun_out/micro_ncu_bit.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python - <<'PY'
import json
for f in ("bench_1gpu", "bench_config4", "bench_config3", "ab_nowidths", "ab_tensorwide", "ab_fused"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["output_sha"], d["config"]["pbs_per_image"], d["check"]["max_abs_deviation_from_clear"], d["roofline"]["frac"], d["roofline_other_pbs_kernel"]["frac"])
    except Exception as e:
        print(f, "missing", e)
PY
ls -la gpurun_out/*.ncu-rep gpurun_out/bench_launches.csv
