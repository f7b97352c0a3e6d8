#!/bin/bash
T=${1:-r02k}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_ -s 60 -c 40 --csv --log-file gpurun_out/${T}_launches_exact.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --exact-edge-scores > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${T}_launches_exact.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value"); u=h.index("Metric Unit")
for r in rows[1:21]: print("%-60s %s %s" % (r[k][:60], r[v], r[u]))
PY
timeout 600 python tools/bench_pipeline.py --pairs 300000 --partners 20 --one-thread-limit 0 2> gpurun_out/${T}_pipe.err | tail -1 > gpurun_out/${T}_bench_pipeline.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_pipeline.json"))
    print("ref", d.get("reference"), "\n  mirror", d.get("mirror_device_ingest"), "\n  breakdown", {k:v for k,v in (d.get("breakdown") or {}).items() if k.startswith("speedup")})
except Exception as ex: print("pipeline failed", ex)
PY
