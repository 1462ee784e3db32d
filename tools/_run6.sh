SGTD_VERIFY_OCC=4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b4.log 2>&1
echo occ=4; grep -o '"value": [0-9.]*' gpurun_out/b4.log | head -1; grep -o '"result_crc": [0-9]*' gpurun_out/b4.log; grep -o '"stage_ms": {[^}]*}' gpurun_out/b4.log
ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/prof_verify4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/pv.log 2>&1
