# multi-GPU bench lines: bash tools/runs/run15.sh N [shards ...]
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1; shift
for S in "$@"; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$S bench.py --gpus $N --shards $S --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_n${N}_s$S.json 2> gpurun_out/r02b_bench_n${N}_s$S.err; echo "N=$N S=$S rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02b_bench_n${N}_s$S.json') if l.startswith('{')][-1])
for k in ("value","ms_per_step","e2e","stage_ms","stage_ms_min_max_over_ranks","result_crc","parity_checked"): print(k, d.get(k))
print(d["config"]["sharding"])
PY
done
