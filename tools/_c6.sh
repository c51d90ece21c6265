mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log)
tail -n 15 gpurun_out/c6_pytest.log | cut -c1-200
(timeout 120 python tools/profile_infer.py 600 > gpurun_out/c6_infer_graph.log 2>&1; timeout 120 python tools/profile_infer.py 600 --no-graph > gpurun_out/c6_infer_nograph.log 2>&1)
grep -h frame gpurun_out/c6_infer_graph.log gpurun_out/c6_infer_nograph.log
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c6_infer_launches.csv python tools/profile_infer.py 600 --no-graph > gpurun_out/c6_infer_ncu.log 2>&1
B="python bench.py --steps 5 --warmup 3 --pretrain 300 --no-fps --no-cpu-baseline --no-vren-ops"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3300 --launch-count 400 --csv --log-file gpurun_out/c6_step_launches.csv $B > gpurun_out/c6_step_ncu.log 2>&1
for k in k_ngp_fwd k_ngp_bwd2 k_grid_scatter_merged; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 290 --launch-count 2 -f -o gpurun_out/r02_$k $B > gpurun_out/c6_ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_infer_march_warp --launch-skip 44 --launch-count 2 -f -o gpurun_out/r02_k_infer_march_warp python tools/profile_infer.py 600 --no-graph > gpurun_out/c6_ncu_infer.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_train_march --launch-skip 290 --launch-count 2 -f -o gpurun_out/r02_k_train_march_c5 $B --workload c5 > gpurun_out/c6_ncu_march_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -n 8
