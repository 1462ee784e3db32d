python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/b.log | head -1; grep -o '"result_crc": [0-9]*' gpurun_out/b.log; grep -o '"stage_ms": {[^}]*}' gpurun_out/b.log
