cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r12_pytest.log
timeout 900 python tools/join_probe.py 100000 1024 "" "collect_unroll=5" "collect_unroll=6" "" > gpurun_out/r12_probe.log 2>&1; cat gpurun_out/r12_probe.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --shards 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r12_bench_n2_s1.json 2> gpurun_out/r12_bench_n2_s1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r12_bench_n2_s1.json') if l.startswith('{')][-1])
for k in ("value","ms_per_step","e2e","stage_ms","result_crc","config","parity_checked","stage_ms_min_max_over_ranks"): print(k, d.get(k))
PY
tail -3 gpurun_out/r12_bench_n2_s1.err
