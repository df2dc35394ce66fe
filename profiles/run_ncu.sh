#!/bin/bash
# profiles/run_ncu.sh -- the ncu passes of /opt/skills/guides/B200_PROFILING.md for bench.py.
# Run under gpurun from the repo root; outputs land in gpurun_out/ and the summaries are copied
# into profiles/ by profiles/summarise.py.  Numbers printed by bench.py under ncu are NOT bench values.
set -x
B="python bench.py --batch 256 --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
# 1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
# 2) full captures of single launches of the kernels we report on
ncu --set full --clock-control none --import-source on -k regex:k_colorspace -c 1 -o gpurun_out/prof_colorspace -f $B > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dwt_cols_t -c 1 -o gpurun_out/prof_dwt_cols -f $B > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_entropy -c 1 -o gpurun_out/prof_entropy -f $B > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_image -s 2 -c 1 -o gpurun_out/prof_ll2_code -f $B > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out
