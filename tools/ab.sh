for m in 0 1; do VIO_SOLVE_TILES=$m python bench.py --steps 30 --warmup 6 --no-cpu > gpurun_out/ab_$m.json 2>gpurun_out/ab_$m.err; tail -2 gpurun_out/ab_$m.err; done
