#!/bin/bash
# N GPUs: peer-memory gather against the NCCL gather (tests, then bench.py with either)
T=${1:-r02w}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -rfs 2>&1 | tail -12 | tee gpurun_out/${T}_pytest_multi_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for X in auto nccl; do
  timeout 600 $TR --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --exchange $X 2> gpurun_out/${T}_bench_${X}.err | tail -1 > gpurun_out/${T}_bench_${N}gpu_${X}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_${N}gpu_${X}.json"))
    print("$X: value %.4e step_ms %.3f | %s" % (d["value"], d["ms_per_step"], d["exchange"][:60]))
except Exception as ex: print("$X: failed", ex)
PY
  tail -2 gpurun_out/${T}_bench_${X}.err
done
