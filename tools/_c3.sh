mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log)
tail -3 gpurun_out/c3_pytest.log
(timeout 120 python tools/profile_infer.py 600 > gpurun_out/c3_infer_graph.log 2>&1; timeout 120 python tools/profile_infer.py 600 --no-graph > gpurun_out/c3_infer_nograph.log 2>&1; timeout 120 python tools/profile_infer.py 600 --unfused > gpurun_out/c3_infer_unfused.log 2>&1)
tail -2 gpurun_out/c3_infer_graph.log gpurun_out/c3_infer_nograph.log gpurun_out/c3_infer_unfused.log
timeout 150 python tools/profile_ref_arm.py fast 300 > gpurun_out/c3_ref_profile.log 2>&1
bash tools/sweep_variants.sh 200 > gpurun_out/c3_sweep.txt 2>&1
cat gpurun_out/c3_sweep.txt
