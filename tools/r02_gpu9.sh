#!/bin/bash
# round 2 evidence pass on one B200: GPU suite, default bench line, launch list, ncu --set full of hc_score_kernel at the
# bench configuration, FNO benchmark + its launch list with DRAM bytes.     tools/r02_gpu9.sh <tag>
T=${1:-r02l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfs > gpurun_out/${T}_pytest_gpu_full.txt 2>&1
grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/${T}_pytest_gpu_full.txt | head -40
timeout 600 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
cut -c1-600 gpurun_out/${T}_bench_1gpu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_ -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_score_kernel -s 2 -c 1 -f -o gpurun_out/${T}_score \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 python tools/bench_fno.py 2> gpurun_out/${T}_fno.err | tail -1 > gpurun_out/${T}_bench_fno.json
cat gpurun_out/${T}_bench_fno.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fno -c 200 --csv \
    --log-file gpurun_out/${T}_fno_launches.csv python tools/bench_fno.py --steps 1 --cpu-edges 0 > /dev/null 2>&1
tail -3 gpurun_out/${T}_fno.err
