#!/bin/bash
# round 2: parity suite + source-level ncu profile of the walk kernel (2 M pairs / 25 M candidates)
T=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_score_kernel -s 3 -c 1 -f -o gpurun_out/${T}_score \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --pairs 2000000 --cands 25000000 > /dev/null 2> gpurun_out/${T}_ncu.err
ncu -i gpurun_out/${T}_score.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/${T}_score_summary.txt
ncu -i gpurun_out/${T}_score.ncu-rep --page source --csv 2>/dev/null > gpurun_out/${T}_score_source.csv
python tools/ncu_source.py 30 < gpurun_out/${T}_score_source.csv > gpurun_out/${T}_score_source_top.txt
head -50 gpurun_out/${T}_score_source_top.txt
timeout 400 python bench.py --no-cpu --steps 5 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); e=d.get('e2e') or {}
print('value %.4e kernel_ms %.3f frac %.3f step_ms %.3f e2e_ms %s results %s' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'], e.get('ms_per_step'), d['results']))"
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/${T}_ab.txt
