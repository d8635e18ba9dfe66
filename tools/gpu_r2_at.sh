#!/bin/bash
# compute-sanitizer over the round's new training kernels (tools/wgrad_sanitize.py: values are asserted inside)
set -u
mkdir -p gpurun_out
timeout -k 5 110 compute-sanitizer --tool racecheck --racecheck-report all python tools/wgrad_sanitize.py > gpurun_out/at_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/at_racecheck.log
timeout -k 5 70 compute-sanitizer --tool memcheck python tools/wgrad_sanitize.py > gpurun_out/at_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/at_memcheck.log
echo done
