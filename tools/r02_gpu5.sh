#!/bin/bash
T=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfs > gpurun_out/${T}_pytest_gpu_full.txt 2>&1
grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/${T}_pytest_gpu_full.txt | head -40
grep -n "Error\|assert " gpurun_out/${T}_pytest_gpu_full.txt | head -20
timeout 600 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_1gpu.json")); e=d.get("e2e") or {}; r=d.get("roofline") or {}
    print("1gpu: value %.4e step_ms %.3f kernel_ms %s frac %s e2e %.4e (%s ms, h2d %s d2h %s) %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms"), r.get("frac"), e.get("value", 0), e.get("ms_per_step"), e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"), e.get("records")))
except Exception as ex: print("failed", ex)
PY
tail -3 gpurun_out/${T}_bench1.err
timeout 600 python tools/bench_pipeline.py --pairs 30000 --partners 20 2> gpurun_out/${T}_pipe_small.err | tail -1 > gpurun_out/${T}_bench_pipeline_small.json
timeout 900 python tools/bench_pipeline.py --pairs 300000 --partners 20 --one-thread-limit 0 2> gpurun_out/${T}_pipe.err | tail -1 > gpurun_out/${T}_bench_pipeline.json
python - <<PY
import json
for f in ("small", ""):
    try:
        d=json.load(open("gpurun_out/${T}_bench_pipeline%s.json" % ("_"+f if f else "")))
        print(f or "full", {k: d.get(k) for k in ("candidates",)}, "ref", d.get("reference"), "\n  mirror", d.get("mirror_device_ingest"), "\n  breakdown", d.get("breakdown"))
    except Exception as ex: print("pipeline failed", ex)
PY
tail -3 gpurun_out/${T}_pipe.err
