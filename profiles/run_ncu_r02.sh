#!/bin/bash
# profiles/run_ncu_r02.sh -- round-2 ncu captures (run under gpurun, one GPU).
#   launch list of the default bench command (cold-cache, serialised: shares only)
#   --set full captures of the kernels the round works on, one launch each, batch 1024
set -x
mkdir -p gpurun_out
B="python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches_r02.log 2>&1
for k in k_front_luma kd_backend kd_inv_rows_t kd_serial_front k_e16_residual k_ll2_code k_entropy k_e20_bands kd_y_markers kd_ll_parallel k_idwt_level k_dwt_level; do
	ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_r02_$k -f $B > gpurun_out/ncu_r02_$k.log 2>&1
done
ls -la gpurun_out | tail -20
