#!/bin/bash
T=${1:-r02t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_small_outputs.py tests/test_gpu_host_mirror.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.txt
timeout 600 python tools/fuzz_parity.py --seeds 24 2>&1 | tail -3 | tee gpurun_out/${T}_fuzz.txt
timeout 600 python bench.py --no-cpu --steps 5 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_1gpu.json")); e=d["e2e"]
print("value %.4e step %.2f e2e %.4e (%.2f ms) pageable %s exact %s" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e.get("pageable", {}).get("ms_per_step"), d["exact_edge_scores"]["ms_per_step"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_exact -c 40 --csv --log-file gpurun_out/${T}_exact_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --exact-edge-scores > /dev/null 2>&1
grep hc_exact gpurun_out/${T}_exact_launches.csv | tail -4
