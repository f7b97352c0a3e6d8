#!/bin/bash
# One GPU-box pass that refreshes the evidence under gpurun_out/ : tests, default bench line, launch list, ncu --set full
# of hc_score_kernel, contig benchmarks, randomised parity sweep.   tools/round_gpu.sh <tag>
T=${1:-rXX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_ -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_score_kernel -s 2 -c 1 -f -o gpurun_out/${T}_score \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 python tools/bench_contigs.py --contigs 20000 --cands 4000000 2>/dev/null | tail -1 > gpurun_out/${T}_bench_contigs.json
timeout 300 python tools/bench_contigs.py --qmax 80 2>/dev/null | tail -1 > gpurun_out/${T}_bench_contigs_packed.json
timeout 900 python tools/fuzz_parity.py --seeds 40 2>&1 | tail -3 | tee gpurun_out/${T}_fuzz.txt
cat gpurun_out/${T}_bench_1gpu.json | cut -c1-400
