#!/bin/bash
# profiles/run_ncu_serial.sh -- full captures (with source) of the latency-bound per-image kernels.
set -x
B2="python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
for k in k_ll2_code k_entropy k_peephole k_e16_residual k_e20_bands; do
	ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof2_$k -f $B2 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
