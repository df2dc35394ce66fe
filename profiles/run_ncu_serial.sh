#!/bin/bash
# profiles/run_ncu_serial.sh -- ncu --set full captures (with source) of the latency-bound encoder stages.
set -x
B="python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_entropy -c 1 -o gpurun_out/prof_entropy -f $B > gpurun_out/ncu_s1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rows -s 3 -c 1 -o gpurun_out/prof_e6d -f $B > gpurun_out/ncu_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_offset_quant -c 1 -o gpurun_out/prof_offq -f $B > gpurun_out/ncu_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ll2_code -c 1 -o gpurun_out/prof_ll2code -f $B > gpurun_out/ncu_s4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_peephole -c 1 -o gpurun_out/prof_peep -f $B > gpurun_out/ncu_s5.log 2>&1
ls -la gpurun_out | head -30
