#!/bin/bash
# round 2, eight-GPU pass: what the host links move at 8 ranks (with / without NUMA binding), then bench.py at N = 8
T=${1:-r02f}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
lscpu | head -30 > gpurun_out/${T}_lscpu.txt 2>&1
numactl -H > gpurun_out/${T}_numa.txt 2>&1 || cat /sys/devices/system/node/node*/cpulist > gpurun_out/${T}_numa.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/bench_hostlink.py 2> gpurun_out/${T}_hostlink.err | tail -1 > gpurun_out/${T}_hostlink.json
timeout 300 $TR --master-port 29522 tools/bench_hostlink.py --numa-bind 2> gpurun_out/${T}_hostlink_numa.err | tail -1 > gpurun_out/${T}_hostlink_numa.json
cat gpurun_out/${T}_hostlink.json gpurun_out/${T}_hostlink_numa.json
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/${T}_bench8.err | tail -1 > gpurun_out/${T}_bench_8gpu.json
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_scale.py::test_store_replicated_on_two_devices -m gpu -q -rfs > gpurun_out/${T}_pytest_multi_gpu.txt 2>&1
tail -5 gpurun_out/${T}_pytest_multi_gpu.txt
for f in 8gpu; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_$f.json")); e=d.get("e2e") or {}
    print("$f: value %.4e step_ms %.3f e2e %.4e (%s ms, h2d %s d2h %s) numa %s" % (d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step"), e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"), d.get("numa")))
except Exception as ex: print("$f: failed", ex)
PY
done
tail -3 gpurun_out/${T}_bench8.err
