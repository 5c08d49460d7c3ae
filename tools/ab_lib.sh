#!/bin/bash
# A/B of alternative builds of the library: bash tools/ab_lib.sh TAG libA.so libB.so ...   (bench.py --no-cpu per library, kernels digest)
TAG=$1; shift
mkdir -p gpurun_out
for L in "$@"; do
  VIO_LIB_NAME=$L timeout 600 python bench.py --steps 30 --warmup 6 --no-cpu > gpurun_out/${TAG}_${L%.so}.json 2> gpurun_out/${TAG}_${L%.so}.err
  python tools/bench_summary.py gpurun_out/${TAG}_${L%.so}.json
done
