#!/bin/bash
# profiles/run_ncu.sh -- the ncu passes of /opt/skills/guides/B200_PROFILING.md for bench.py.
# Run under gpurun from the repo root; outputs land in gpurun_out/ and the summaries are written
# into profiles/ by profiles/summarise.py.  Numbers printed by bench.py under ncu are NOT bench values.
set -x
B="python bench.py --batch 256 --steps 1 --warmup 3 --no-cpu-baseline"
B2="python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
# 1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
# 2) full captures (with source) of single launches of the kernels we report on
ncu --set full --clock-control none --import-source on -k regex:k_front_luma -s 1 -c 1 -o gpurun_out/prof_front_luma -f $B2 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dwt_level -s 4 -c 3 -o gpurun_out/prof_dwt_level -f $B2 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_entropy -c 1 -o gpurun_out/prof_entropy -f $B2 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ll2_code -c 1 -o gpurun_out/prof_ll2_code -f $B2 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_y_quant_scan -c 1 -o gpurun_out/prof_quant_scan -f $B2 > gpurun_out/ncu_full5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kd_serial_front -c 1 -o gpurun_out/prof_dec_front -f $B2 > gpurun_out/ncu_full6.log 2>&1
ls -la gpurun_out
