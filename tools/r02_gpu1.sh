#!/bin/bash
# round 2, first GPU pass: anchor walk correctness, parity suite, A/B timing, L1 metrics
T=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${T}_gpu.txt
timeout 900 python tools/check_walk.py > gpurun_out/${T}_check_walk.txt 2>&1; echo "check_walk rc=$?" | tee -a gpurun_out/${T}_check_walk.txt
tail -5 gpurun_out/${T}_check_walk.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_gpu.txt
for mode in walk nowalk; do
  if [ $mode == nowalk ]; then export HC_NO_ANCHOR_WALK=1; else unset HC_NO_ANCHOR_WALK; fi
  timeout 400 python bench.py --no-cpu --steps 5 2> gpurun_out/${T}_bench_${mode}.err | tail -1 > gpurun_out/${T}_bench_${mode}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${mode}.json")); e=d.get("e2e") or {}
    print("${mode}: value %.4e kernel_ms %.3f frac %.3f step_ms %.3f e2e_ms %s results %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["ms_per_step"], e.get("ms_per_step"), d["results"]))
except Exception as ex: print("${mode}: failed", ex)
PY
done
unset HC_NO_ANCHOR_WALK
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/${T}_ab.txt
bash tools/ab_ncu.sh > gpurun_out/${T}_ab_ncu.txt 2>&1
tail -60 gpurun_out/${T}_ab_ncu.txt
