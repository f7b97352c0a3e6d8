#!/bin/bash
# One GPU-box pass that refreshes the evidence under gpurun_out/ (copy what is to be kept into profiles/): full GPU suite,
# memcheck over the host-buffer paths, default bench line, launch list, ncu --set full of hc_score_kernel, fuzz, FNO.   tools/round_gpu.sh <tag>
T=${1:-rXX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfs > gpurun_out/${T}_pytest_gpu_full.txt 2>&1
grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/${T}_pytest_gpu_full.txt | head -20
HC_STAGE_CHUNK=65536 HC_STAGE_MIN=4096 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_adjacency.py tests/test_gpu_fno.py tests/test_gpu_fastq.py tests/test_gpu_stage.py tests/test_gpu_dedup.py tests/test_gpu_ingest.py -m gpu -q -x -k "not random_multigraph" 2>&1 | tail -4 | tee gpurun_out/${T}_memcheck.txt
timeout 900 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_1gpu.json")); e=d["e2e"]
print("value %.4e step %.2f kernel %.2f frac %.3f e2e %.4e (%.2f ms) pageable %s exact %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], e["value"], e["ms_per_step"], e.get("pageable", {}).get("ms_per_step"), d["exact_edge_scores"]["ms_per_step"]))
print(d.get("files_to_graph")); print(d["roofline"]["traffic_source"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_ -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_score_kernel -s 2 -c 1 -f -o gpurun_out/${T}_score \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 600 python tools/fuzz_parity.py --seeds 40 2>&1 | tail -2 | tee gpurun_out/${T}_fuzz.txt
timeout 300 python tools/bench_fno.py 2>/dev/null | tail -1 > gpurun_out/${T}_bench_fno.json; cut -c1-250 gpurun_out/${T}_bench_fno.json
