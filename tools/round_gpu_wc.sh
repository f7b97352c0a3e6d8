#!/bin/bash
# Eight GPUs: do write-combined source buffers change what the host links move?  Host-link benchmark with and without, then
# bench.py's e2e leg with the candidate records in write-combined memory (HC_BENCH_WC=1).   gpurun --gpus 8 -- bash tools/round_gpu_wc.sh <tag>
T=${1:-rXX}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29551 tools/bench_hostlink.py --mb-in 720 --mb-out 180 --write-combined 2> gpurun_out/${T}_hostlink_wc.err | tail -1 > gpurun_out/${T}_hostlink_wc.json
cat gpurun_out/${T}_hostlink_wc.json
HC_BENCH_WC=1 timeout 600 $TR --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/${T}_bench8_wc.err | tail -1 > gpurun_out/${T}_bench_8gpu_wc.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_8gpu_wc.json")); e=d.get("e2e") or {}
    print("8gpu wc: value %.4e step_ms %.3f e2e %.4e (%s ms)" % (d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step")))
except Exception as ex: print("failed", ex)
PY
tail -2 gpurun_out/${T}_bench8_wc.err
