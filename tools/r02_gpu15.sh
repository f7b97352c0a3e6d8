#!/bin/bash
# round 2, eight-GPU pass with the 6-byte records: multi-GPU tests (log kept), host links at the e2e sizes, bench.py at N = 8
T=${1:-r02u}
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/${T}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_scale.py::test_store_replicated_on_two_devices -m gpu -q -rfs > gpurun_out/${T}_pytest_multi_gpu.txt 2>&1
tail -5 gpurun_out/${T}_pytest_multi_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/bench_hostlink.py --mb-in 720 --mb-out 180 2> gpurun_out/${T}_hostlink.err | tail -1 > gpurun_out/${T}_hostlink.json
cat gpurun_out/${T}_hostlink.json
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/${T}_bench8.err | tail -1 > gpurun_out/${T}_bench_8gpu.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_8gpu.json")); e=d.get("e2e") or {}
    print("8gpu: value %.4e step_ms %.3f e2e %.4e (%s ms, h2d %s d2h %s) %s" % (d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step"), e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"), e.get("records")))
except Exception as ex: print("8gpu: failed", ex)
PY
tail -3 gpurun_out/${T}_bench8.err
