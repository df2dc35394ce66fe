#!/bin/bash
# profiles/run_ncu_dec.sh -- ncu --set full captures (with source) of the decoder's serial stages.
set -x
B="python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:kd_image_lut -c 1 -o gpurun_out/prof_dprefix -f $B > gpurun_out/ncu_d1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kd_image -s 2 -c 1 -o gpurun_out/prof_dmarkers -f $B > gpurun_out/ncu_d2.log 2>&1
ls -la gpurun_out | tail -5
