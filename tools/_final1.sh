set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/bench_r01f_n1.json 2> gpurun_out/bench_r01f_n1.err
python bench.py --workload seq --scans 1024 --queries 256 --batch 128 --no-cpu-baseline > gpurun_out/bench_r01f_seq128.json 2>/dev/null
python bench.py --workload seq --scans 1024 --queries 512 --batch 512 --no-cpu-baseline > gpurun_out/bench_r01f_seq512.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_vote_join|k_collect_inv|k_verify|k_query_index|k_hypotheses|k_probe_emit" -s 6 -c 6 -o gpurun_out/prof_r01f -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_bench_f.log 2>&1
cat gpurun_out/final_tests.log; tail -1 gpurun_out/final_smoke.log; cut -c1-300 gpurun_out/bench_r01f_n1.json
