#!/bin/bash
T=${1:-r02g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rfs > gpurun_out/${T}_pytest_gpu_full.txt 2>&1
grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/${T}_pytest_gpu_full.txt | head -40
grep -n "Error\|assert " gpurun_out/${T}_pytest_gpu_full.txt | head -30
