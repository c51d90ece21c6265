mkdir -p gpurun_out
(NGP_BWD_VARIANT=2 timeout 150 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/c10_pytest_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest_v2.log)
grep -E "passed|failed|FAILED|rror|rc=" gpurun_out/c10_pytest_v2.log | tail -n 5 | cut -c1-220
(NGP_BWD_VARIANT=2 timeout 200 python bench.py --steps 200 --warmup 5 --no-fps --no-cpu-baseline --no-vren-ops > gpurun_out/c10_bench_v2.log 2>&1; echo "rc=$?" >> gpurun_out/c10_bench_v2.log)
python - <<'P'
import json
for line in open("gpurun_out/c10_bench_v2.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("variant 2 step %.4f ms" % d["ms_per_step"], {k["kernel"]: (round(k["ms_per_launch"] * 1e3, 1), round(k["ms_per_launch_cold_l2"] * 1e3, 1)) for k in d["roofline"]["kernels"]})
P
NGP_BWD_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ngp_bwd3 --launch-skip 1100 --launch-count 1 -f -o gpurun_out/r02_k_ngp_bwd3 python bench.py --steps 5 --warmup 3 --pretrain 1100 --no-fps --no-cpu-baseline --no-vren-ops > gpurun_out/c10_ncu_bwd3.log 2>&1
ls -la gpurun_out/r02_k_ngp_bwd3.ncu-rep
