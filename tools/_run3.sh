python -m pytest tests/test_gpu_parity.py -x -q -k "long_match" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:"k_collect_inv|k_query_index" -s 2 -c 2 -o gpurun_out/prof_collect2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/pc.log 2>&1
tail -2 gpurun_out/pc.log | cut -c1-150
