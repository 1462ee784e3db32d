ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_e.log 2>&1
tail -1 gpurun_out/launches_bench_e.log | cut -c1-200
