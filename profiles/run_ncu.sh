#!/bin/bash
# profiles/run_ncu.sh -- the ncu passes of /opt/skills/guides/B200_PROFILING.md for bench.py.
# Run under gpurun from the repo root; outputs land in gpurun_out/ and the summaries are written
# into profiles/ by profiles/summarise.py.  Numbers printed by bench.py under ncu are NOT bench values.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"   # the bench workload itself (batch 4096), fewer timed steps
B2="python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
# 1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
# 2) full captures (with source) of single launches of the kernels we report on
for k in k_front_luma k_entropy k_ll2_code k_y_quant_scan k_patterns k_e20_bands k_e16_residual k_peephole kd_serial_front; do
	ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_$k -f $B2 > gpurun_out/ncu_full_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_dwt_level -s 4 -c 3 -o gpurun_out/prof_k_dwt_level -f $B2 > gpurun_out/ncu_full_dwt.log 2>&1
ls -la gpurun_out
