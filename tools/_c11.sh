mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log)
grep -E "passed|failed|FAILED|rror|rc=" gpurun_out/c11_pytest.log | tail -n 5 | cut -c1-220
(timeout 240 python bench.py --steps 200 --warmup 5 > gpurun_out/c11_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/c11_bench_n1.log)
python - <<'P'
import json
for line in open("gpurun_out/c11_bench_n1.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("step %.4f ms  value %.2fM e2e %.2fM fps %.1f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["render_fps"]["value"]), {k["kernel"]: (round(k["ms_per_launch"] * 1e3, 1), round(k["ms_per_launch_cold_l2"] * 1e3, 1)) for k in d["roofline"]["kernels"]})
P
