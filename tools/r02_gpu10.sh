#!/bin/bash
# FNO rework: tests, benchmark with phase times, launch list with DRAM bytes, memcheck.   tools/r02_gpu10.sh <tag>
T=${1:-r02m}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fno.py tests/test_gpu_fno_host.py tests/test_gpu_stage.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_fno.txt
for th in 8 16 24; do HC_STAGE_THREADS=$th timeout 300 python tools/bench_fno.py --steps 5 --cpu-edges 0 2>/dev/null | tail -1 | cut -c1-330; done
timeout 300 python tools/bench_fno.py --steps 5 2>/dev/null | tail -1 > gpurun_out/${T}_bench_fno.json
HC_FNO_TIMING=1 timeout 300 python tools/bench_fno.py --steps 2 --cpu-edges 0 2> gpurun_out/${T}_fno.err > /dev/null
cat gpurun_out/${T}_bench_fno.json; tail -45 gpurun_out/${T}_fno.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fno|ffw" -c 300 --csv \
    --log-file gpurun_out/${T}_fno_launches.csv python tools/bench_fno.py --steps 1 --cpu-edges 0 > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fno.py -m gpu -q -x -k "partitions or beyond or small" 2>&1 | tail -5 | tee gpurun_out/${T}_memcheck_fno.txt
