#!/bin/bash
# Round-2 GPU visit: ncu launch list restricted to the library's kernels, ncu --set full captures of the heavy kernels (steady state),
# all from `bench.py --no-cpu --no-extras` (the command whose kernels the bench times).  Usage: gpurun -- 'bash tools/gpu_round2.sh r02x'
TAG=${1:-r02x}
ONLY=${2:-all}          # all | quick (launch list + solve capture only)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-extras"
# launch list: every launch of the library's kernels (names live in namespaces fe:: / be:: or end in _kernel), prologue included
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(pyr_down|lk|post_track|eig_candidates|select|clahe_lut|clahe_apply|imu|addfeat|triangulate|prepare|solve|post_solve|marg|finish|clear_init_pending|init_state|set_init)_kernel' -c 1500 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
  $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^solve_kernel' --launch-skip 12 --launch-count 1 \
  -o gpurun_out/${TAG}_solve -f $B > gpurun_out/${TAG}_ncu_solve.log 2>&1
if [ "$ONLY" = "all" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^marg_kernel' --launch-skip 12 --launch-count 1 \
  -o gpurun_out/${TAG}_marg -f $B > gpurun_out/${TAG}_ncu_marg.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lk_kernel|eig_candidates_kernel|pyr_down_kernel|post_track_kernel|select_kernel' \
  --launch-skip 140 --launch-count 8 -o gpurun_out/${TAG}_fe -f $B > gpurun_out/${TAG}_ncu_fe.log 2>&1
fi
ls -la gpurun_out | grep ${TAG}
