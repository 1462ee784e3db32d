# final single-GPU refresh: fast tests, bench line, ncu capture (traffic record), launch list
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_instances.py -x -q > gpurun_out/r23_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r23_pytest.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vote_join|k_verify|k_collect_inv|k_probe_emit|k_query_index|k_topk|k_hypotheses" --launch-skip 7 --launch-count 7 -o gpurun_out/r02b_search python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu_search.log 2>&1; echo "ncu search rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_100k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_launches_bench.log 2>&1; echo "launch list rc=$?"
