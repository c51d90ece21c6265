mkdir -p gpurun_out
(NGP_BWD_VARIANT=2 timeout 150 python -m pytest tests/test_network_gpu.py tests/test_fused_gpu.py tests/test_round2_gpu.py -x -q > gpurun_out/c9_pytest_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest_v2.log)
grep -E "passed|failed|FAILED|rror|rc=" gpurun_out/c9_pytest_v2.log | tail -n 8 | cut -c1-220
for v in 1 2; do
  (NGP_BWD_VARIANT=$v timeout 200 python bench.py --steps 200 --warmup 5 --no-fps --no-cpu-baseline --no-vren-ops > gpurun_out/c9_bench_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/c9_bench_v$v.log)
done
python - <<'P'
import json
for v in (1, 2):
    for line in open("gpurun_out/c9_bench_v%d.log" % v):
        if line.startswith("{"):
            d = json.loads(line)
            print("variant", v, "step %.4f ms" % d["ms_per_step"], "psnr %.2f" % d["config"]["train_psnr_last_batch"],
                  {k["kernel"]: (round(k["ms_per_launch"] * 1e3, 1), round(k["ms_per_launch_cold_l2"] * 1e3, 1)) for k in d["roofline"]["kernels"]})
            break
    else:
        print("variant", v, "no JSON line"); print(open("gpurun_out/c9_bench_v%d.log" % v).read()[-1500:])
P
