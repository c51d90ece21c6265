mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log)
grep -E "passed|failed|FAILED|Error" gpurun_out/c7_pytest.log | tail -n 8 | cut -c1-200
timeout 120 python tools/profile_infer.py 600 > gpurun_out/c7_infer_graph.log 2>&1; grep -h frame gpurun_out/c7_infer_graph.log
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_infer_launches.csv python tools/profile_infer.py 600 --no-graph > gpurun_out/c7_infer_ncu.log 2>&1
(timeout 240 python bench.py --steps 200 --warmup 5 > gpurun_out/c7_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/c7_bench_n1.log)
B="python bench.py --steps 30 --warmup 5 --no-fps --no-cpu-baseline --no-vren-ops"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_step_launches.csv $B > gpurun_out/c7_step_ncu.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_step_launches_c5.csv $B --workload c5 > gpurun_out/c7_step_ncu_c5.log 2>&1
grep -h "rc=" gpurun_out/c7_*.log
