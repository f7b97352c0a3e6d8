#!/bin/bash
# Eight GPUs of one box: bench.py with the peer-memory gather (default) and with NCCL, the multi-GPU tests.   gpurun --gpus 8 -- bash tools/round_gpu_multi.sh <tag>
T=${1:-rXX}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/${T}_bench8.err | tail -1 > gpurun_out/${T}_bench_8gpu.json
timeout 600 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --exchange nccl 2> gpurun_out/${T}_bench8_nccl.err | tail -1 > gpurun_out/${T}_bench_8gpu_nccl.json
python - <<PY
import json
for f in ("", "_nccl"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_8gpu%s.json" % f)); e=d.get("e2e") or {}
        print("8gpu%s: value %.4e step_ms %.3f e2e %.4e (%s ms) | %s" % (f, d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step"), d["exchange"][:50]))
    except Exception as ex: print("8gpu%s: failed" % f, ex)
PY
grep "gather:" gpurun_out/${T}_bench8.err | head -2; tail -2 gpurun_out/${T}_bench8.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_multi_gpu.txt
