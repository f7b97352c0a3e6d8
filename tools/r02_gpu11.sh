#!/bin/bash
# mirror: tests, files-in -> graph-out at 300 k and 3 M pairs with the step timers; FNO bench with staging-thread sweep
T=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_mirror.py tests/test_gpu_fno_host.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_mirror.txt
HC_MIRROR_TIMING=1 timeout 900 python tools/bench_pipeline.py --pairs 300000 --partners 20 --one-thread-limit 0 2> gpurun_out/${T}_pipe.err | tail -1 > gpurun_out/${T}_bench_pipeline.json
HC_MIRROR_TIMING=1 timeout 1500 python tools/bench_pipeline.py --pairs 3000000 --partners 20 --one-thread-limit 0 2> gpurun_out/${T}_pipe_3m.err | tail -1 > gpurun_out/${T}_bench_pipeline_3m.json
python - <<PY
import json
for f in ("", "_3m"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_pipeline%s.json" % f))
        print(f or "300k", d.get("candidates"), "ref", d.get("reference"), "\n  host", d.get("mirror_host_parsers"), "\n  mirror", d.get("mirror_device_ingest"), "\n  breakdown", {k: v for k, v in d.get("breakdown", {}).items() if k.startswith("speedup")})
    except Exception as ex: print("pipeline failed", ex)
PY
tail -32 gpurun_out/${T}_pipe.err; tail -16 gpurun_out/${T}_pipe_3m.err
for th in 8 16 24; do HC_STAGE_THREADS=$th timeout 300 python tools/bench_fno.py --steps 5 --cpu-edges 0 2>/dev/null | tail -1 | cut -c1-330; done
timeout 300 python tools/bench_fno.py --steps 5 2>/dev/null | tail -1 > gpurun_out/${T}_bench_fno.json
