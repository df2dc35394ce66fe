#!/bin/bash
# profiles/run_ncu_front.sh -- ncu --set full captures of the fused front-end kernels (one launch each).
set -x
B="python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_front_luma -s 1 -c 1 -o gpurun_out/prof_front_luma -f $B > gpurun_out/ncu_front1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dwt_level -s 4 -c 3 -o gpurun_out/prof_dwt_level -f $B > gpurun_out/ncu_front2.log 2>&1
ls -la gpurun_out
