# final single-GPU evidence of the round: bench line, launch lists, ncu --set full captures
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --workload seq --steps 4 --warmup 2 > gpurun_out/r02b_seq.json 2> gpurun_out/r02b_seq.err; echo "seq rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_100k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vote_join|k_verify|k_collect_inv|k_probe_emit|k_query_index|k_topk|k_hypotheses" --launch-skip 7 --launch-count 7 -o gpurun_out/r02b_search python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu_search.log 2>&1; echo "ncu search rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_s1|k_dcvc|k_fill|Radix|radix" --csv --log-file gpurun_out/r02b_launches_stage1.csv python tools/s1_probe.py 128 2 > /dev/null 2>&1; echo "s1 launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dcvc_replay_cc|k_dcvc_prepare|k_dcvc_rows|k_s1_lcentroid|k_s1_finish|k_s1_gather" --launch-skip 8 --launch-count 8 -o gpurun_out/r02b_stage1 python tools/s1_probe.py 128 2 > gpurun_out/r02b_ncu_stage1.log 2>&1; echo "ncu s1 rc=$?"
S1_TRACE=2 timeout 600 python tools/s1_probe.py 128 5 > gpurun_out/r02b_s1_trace.log 2>&1; grep "rep " gpurun_out/r02b_s1_trace.log
ls -la gpurun_out | tail -20
