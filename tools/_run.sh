python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2 > gpurun_out/t.log; python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2>&1; cat gpurun_out/t.log; python -c '
import json
for l in open("gpurun_out/b.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["config"].get("result_crc"), d["config"].get("stage_ms"))
'
