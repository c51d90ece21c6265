#!/bin/bash
# A/B of the env-switched kernel / schedule variants through bench.py (1 GPU): step time + the per-kernel timings of its
# roofline block, one JSON line per variant in gpurun_out/sweep_<tag>.log.   bash tools/sweep_variants.sh [steps]
STEPS=${1:-200}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 240 python bench.py --steps $STEPS --warmup 5 --no-fps --no-cpu-baseline --no-vren-ops > gpurun_out/sweep_$tag.log 2>&1
  python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
for line in open("gpurun_out/sweep_%s.log" % tag):
    if line.startswith("{"):
        d = json.loads(line)
        ks = {k["kernel"]: (round(k["ms_per_launch"] * 1e3, 1), round(k["ms_per_launch_cold_l2"] * 1e3, 1)) for k in d["roofline"]["kernels"]}
        print("%-28s step %.4f ms  e2e %.2f M  psnr %.2f  kernels(warm,cold us) %s" % (tag, d["ms_per_step"], d["e2e"]["value"] / 1e6,
              d["config"]["train_psnr_last_batch"], ks), flush=True)
        break
else:
    print(tag, "FAILED", flush=True)
P
}
run base NGP_X=0
run prio0 NGP_PRIORITY=0
run paired NGP_GATHER_PAIRED=1
run depth2 NGP_GATHER_DEPTH=2
run paired_depth2 NGP_GATHER_PAIRED=1 NGP_GATHER_DEPTH=2
run v1_depth2 NGP_FWD_VARIANT=1 NGP_GATHER_DEPTH=2
run v1_paired_depth2 NGP_FWD_VARIANT=1 NGP_GATHER_PAIRED=1 NGP_GATHER_DEPTH=2
run scat_paired NGP_SCATTER_PAIRED=1
run all NGP_GATHER_PAIRED=1 NGP_GATHER_DEPTH=2 NGP_SCATTER_PAIRED=1
