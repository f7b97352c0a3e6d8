#!/bin/bash
T=${1:-r02z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfs > gpurun_out/${T}_pytest_gpu_full.txt 2>&1
grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/${T}_pytest_gpu_full.txt | head -20
HC_STAGE_CHUNK=65536 HC_STAGE_MIN=4096 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_adjacency.py tests/test_gpu_fno.py tests/test_gpu_fastq.py tests/test_gpu_stage.py tests/test_gpu_dedup.py tests/test_gpu_ingest.py -m gpu -q -x -k "not random_multigraph" 2>&1 | tail -4 | tee gpurun_out/${T}_memcheck.txt
