#!/bin/bash
T=${1:-r02v}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_adjacency.py tests/test_gpu_fastq.py tests/test_gpu_host_mirror.py tests/test_gpu_stage.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.txt
timeout 300 python tools/bench_adjacency.py 2>&1 | tail -1 | tee gpurun_out/${T}_bench_adjacency.json
HC_MIRROR_TIMING=1 timeout 900 python tools/bench_pipeline.py --pairs 300000 --partners 20 --one-thread-limit 0 --skip-host-parsers 2> gpurun_out/${T}_pipe.err | tail -1 > gpurun_out/${T}_bench_pipeline.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_pipeline.json"))
print(d.get("candidates"), "ref", d.get("reference"), "\n  mirror", d.get("mirror_device_ingest"), "\n  breakdown", {k: v for k, v in d.get("breakdown", {}).items() if k.startswith("speedup")})
PY
