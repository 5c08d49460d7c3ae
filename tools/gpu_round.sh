#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh TAG [tests|notests] [full|nofull]
TAG=${1:-rXX}
DO_TESTS=${2:-tests}
DO_FULL=${3:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "$DO_TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 900 python bench.py --steps 30 --warmup 6 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
if [ "$DO_FULL" = "full" ]; then
  # back end: skip the prologue (first solves), capture one steady-state launch of each heavy kernel
  # (the first ten launches of each are no-ops while the window fills; tools/ncu_summary.py lists every captured launch)
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:(^|:)solve_kernel' --launch-skip 12 --launch-count 1 \
    -o gpurun_out/${TAG}_be -f python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_be.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'marg_kernel' --launch-skip 12 --launch-count 1 \
    -o gpurun_out/${TAG}_marg -f python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_marg.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lk_kernel|eig_candidates_kernel|pyr_down_kernel|post_track_kernel|select_kernel' \
    --launch-skip 140 --launch-count 8 -o gpurun_out/${TAG}_fe -f python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_fe.log 2>&1
fi
ls -la gpurun_out | tail -20
