mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/check_p2p.py p2p > gpurun_out/c5_check_p2p_n8.log 2>&1; echo "rc=$?" >> gpurun_out/c5_check_p2p_n8.log
timeout 200 $TR --master-port 29512 tools/check_p2p.py nvls > gpurun_out/c5_check_nvls_n8.log 2>&1; echo "rc=$?" >> gpurun_out/c5_check_nvls_n8.log
timeout 300 $TR --master-port 29513 bench.py --gpus 8 --steps 200 --warmup 5 --ddp p2p > gpurun_out/c5_bench_n8_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c5_bench_n8_p2p.log
timeout 300 $TR --master-port 29514 bench.py --gpus 8 --steps 200 --warmup 5 --ddp nvls --no-fps --no-vren-ops > gpurun_out/c5_bench_n8_nvls.log 2>&1; echo "rc=$?" >> gpurun_out/c5_bench_n8_nvls.log
timeout 300 $TR --master-port 29515 bench.py --gpus 8 --steps 200 --warmup 5 --workload c5 --no-vren-ops > gpurun_out/c5_bench_n8_c5.log 2>&1; echo "rc=$?" >> gpurun_out/c5_bench_n8_c5.log
grep -h "rc=" gpurun_out/c5_*.log
grep -h "back to back" gpurun_out/c5_check_*.log | head -4 | cut -c1-200
