#!/bin/bash
# renderer launch list + ncu capture of the fused MLP kernel; compute-sanitizer over the tests of the new TMA kernels
set -u
mkdir -p gpurun_out
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/z_render_launches.csv python tools/run_render_branch.py > gpurun_out/z_render_ncu.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:render_mlp_kernel -c 1 -o gpurun_out/z_render_mlp python tools/run_render_branch.py > gpurun_out/z_render_mlp_ncu.log 2>&1
K="s2_slab or transpose3d_slab or o_slab or s2k5 or conv2d_slab or conv3d_slab_bf16x3 or cuda_graph"
timeout -k 10 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "$K" > gpurun_out/z_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/z_memcheck.log
timeout -k 10 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "s2_slab or transpose3d_slab or o_slab or s2k5" > gpurun_out/z_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/z_racecheck.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" > gpurun_out/z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_smoke.log
echo done
