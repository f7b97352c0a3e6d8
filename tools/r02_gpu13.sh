#!/bin/bash
# kernel A/B (variants under haploconduct_b200/lib/variants), ncu --set full of the many-candidate hc_exact_kernel, default bench line
T=${1:-r02s}
mkdir -p gpurun_out
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/${T}_ab.txt
bash tools/ab_bench.sh --e2e --e2e-no-output 2>&1 | tee -a gpurun_out/${T}_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_exact_kernel -s 2 -c 1 -f -o gpurun_out/${T}_exact \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --exact-edge-scores > /dev/null 2>&1
ls -la gpurun_out/${T}_exact.ncu-rep
timeout 900 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_1gpu.json")); e=d["e2e"]
print("value %.4e step %.2f e2e %.4e (%.2f ms) pageable %s exact %s" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e.get("pageable", {}).get("ms_per_step"), d["exact_edge_scores"]["ms_per_step"]))
print(d.get("files_to_graph"))
PY
