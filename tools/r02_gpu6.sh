#!/bin/bash
T=${1:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_small_outputs.py tests/test_gpu_host_mirror.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_1gpu.json")); e=d.get("e2e") or {}; r=d.get("roofline") or {}
    print("1gpu: value %.4e step_ms %.3f kernel_ms %s frac %s e2e %.4e (%s ms) exact %s pageable %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms"), r.get("frac"), e.get("value", 0), e.get("ms_per_step"), d.get("exact_edge_scores"), e.get("pageable")))
except Exception as ex: print("failed", ex)
PY
tail -3 gpurun_out/${T}_bench1.err
timeout 900 python tools/bench_pipeline.py --pairs 300000 --partners 20 --one-thread-limit 0 2> gpurun_out/${T}_pipe.err | tail -1 > gpurun_out/${T}_bench_pipeline.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_pipeline.json"))
    print("ref", d.get("reference"), "\n  mirror", d.get("mirror_device_ingest"), "\n  breakdown", d.get("breakdown"))
except Exception as ex: print("pipeline failed", ex)
PY
