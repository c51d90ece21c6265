mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log)
tail -n 3 gpurun_out/c4_pytest.log
(timeout 120 python tools/profile_infer.py 600 > gpurun_out/c4_infer_graph.log 2>&1; timeout 120 python tools/profile_infer.py 600 --no-graph > gpurun_out/c4_infer_nograph.log 2>&1)
grep -h frame gpurun_out/c4_infer_graph.log gpurun_out/c4_infer_nograph.log
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_infer_launches.csv python tools/profile_infer.py 600 --no-graph > gpurun_out/c4_infer_ncu.log 2>&1
(timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c4_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/c4_bench_ref.log)
(timeout 200 python bench.py --steps 200 --warmup 5 > gpurun_out/c4_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/c4_bench_n1.log)
(timeout 240 python bench.py --impl reference --workload c5 --steps 20 --warmup 5 > gpurun_out/c4_bench_ref_c5.log 2>&1; echo "rc=$?" >> gpurun_out/c4_bench_ref_c5.log)
(timeout 200 python bench.py --workload c5 --steps 200 --warmup 5 > gpurun_out/c4_bench_c5.log 2>&1; echo "rc=$?" >> gpurun_out/c4_bench_c5.log)
(timeout 400 python tools/psnr_parity.py 10000 gpurun_out/c4_psnr_parity.json > gpurun_out/c4_psnr.log 2>&1; echo "rc=$?" >> gpurun_out/c4_psnr.log)
grep -h "step\|rc=" gpurun_out/c4_psnr.log | tail -n 12
grep -h "rc=" gpurun_out/c4_bench*.log
