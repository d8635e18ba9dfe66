#!/bin/bash
# round-2 GPU call A: full GPU test suite, default bench, sanitizer subset, launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x \
  -k "test_cascade_eval_matches_reference_golden or test_conv3d_tensor_core_vs_oracle or test_conv_transpose3d_tensor_core or test_unsup_loss_value or test_geo_consistency_matches or test_costvolume_variance_golden or test_conv3d_gather_gemm" \
  > gpurun_out/a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/a_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -m gpu -q -x \
  -k "(test_conv3d_tensor_core_vs_oracle and 3-8-8) or (test_conv3d_gather_gemm and 3-16-16) or (test_conv_transpose3d_tensor_core and 3-16-8) or test_unsup_loss_value" \
  > gpurun_out/a_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/a_racecheck.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines > gpurun_out/a_bench_under_ncu.log 2>&1
echo done
