python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
for occ in 5 6; do
SGTD_VERIFY_OCC=$occ python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b$occ.log 2>&1
echo occ=$occ; grep -o '"value": [0-9.]*' gpurun_out/b$occ.log | head -1; grep -o '"result_crc": [0-9]*' gpurun_out/b$occ.log; grep -o '"stage_ms": {[^}]*}' gpurun_out/b$occ.log
done
