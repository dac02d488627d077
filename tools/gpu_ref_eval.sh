#!/bin/bash
# The reference's UNMODIFIED homomorphic_eval.py (staged, git-ignored, under baseline/_ref) on this backend, fhe_mode=execute, on a B200.
mkdir -p gpurun_out
timeout 1500 python tools/run_reference_eval.py --reference baseline/_ref/dct-cryptonets --workdir /tmp/ref_eval --synthetic-cifar 200 -- \
  --dataset cifar10 --model ResNet20qat --dct_status --channels 24 --filter_size 4 --image_size_dct 16 --bit_width 4 \
  --fhe_mode execute --calib_batch_size 100 --test_batch_size 1 --test_subset 1 --rounding_threshold_bits 6 --n_bits 5 --p_error 0.01 \
  > gpurun_out/reference_eval_execute.log 2>&1
echo rc=$?
grep -v "^\s*$" gpurun_out/reference_eval_execute.log | grep -v "%|" | tail -40 | cut -c1-220
