#!/bin/bash
# end-of-round validation: full GPU suite, smoke, default bench (all arms), reference arm, final launch lists
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/ac_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ac_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ac_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ac_smoke.log
timeout -k 10 1200 python bench.py > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; echo "bench rc=$?" >> gpurun_out/ac_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/ac_bench_ref.json 2> gpurun_out/ac_bench_ref.err; echo "rc=$?" >> gpurun_out/ac_bench_ref.err
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ac_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --no-graphs --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines > gpurun_out/ac_bench_under_ncu.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ac_render_launches.csv python tools/run_render_branch.py > gpurun_out/ac_render_ncu.log 2>&1
echo done
