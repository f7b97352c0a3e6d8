#!/bin/bash
# mirror after the staged copies: tests (mirror, ingest, stage, small outputs, parity on pageable buffers), pipeline at 300 k / 3 M pairs, default bench
T=${1:-r02p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_mirror.txt
HC_MIRROR_TIMING=1 timeout 900 python tools/bench_pipeline.py --pairs 300000 --partners 20 --one-thread-limit 0 2> gpurun_out/${T}_pipe.err | tail -1 > gpurun_out/${T}_bench_pipeline.json
HC_MIRROR_TIMING=1 timeout 1500 python tools/bench_pipeline.py --pairs 3000000 --partners 20 --one-thread-limit 0 --skip-host-parsers 2> gpurun_out/${T}_pipe_3m.err | tail -1 > gpurun_out/${T}_bench_pipeline_3m.json
python - <<PY
import json
for f in ("", "_3m"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_pipeline%s.json" % f))
        print(f or "300k", d.get("candidates"), "ref", d.get("reference"), "\n  mirror", d.get("mirror_device_ingest"), "\n  breakdown", {k: v for k, v in d.get("breakdown", {}).items() if k.startswith("speedup")})
    except Exception as ex: print("pipeline failed", ex)
PY
tail -10 gpurun_out/${T}_pipe.err; tail -10 gpurun_out/${T}_pipe_3m.err
timeout 600 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_1gpu.json")); e=d["e2e"]
print("value %.4e step %.2f e2e %.4e (%.2f ms) pageable %s" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e.get("pageable")))
PY
