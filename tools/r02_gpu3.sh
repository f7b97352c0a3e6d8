#!/bin/bash
# round 2, two-GPU pass: the whole GPU suite (multi-GPU tests included), bench at N = 1 and N = 2, the reference arm
T=${1:-r02e}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${T}_gpus.txt
timeout 1200 python -m pytest tests -m gpu -q -rs 2>&1 | tail -25 | tee gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py 2> gpurun_out/${T}_bench1.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
    2> gpurun_out/${T}_bench2.err | tail -1 > gpurun_out/${T}_bench_2gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${T}_ref.err | tail -1 > gpurun_out/${T}_bench_ref.json
for f in 1gpu 2gpu ref; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_$f.json")); e=d.get("e2e") or {}; r=d.get("roofline") or {}
    print("$f: value %.4e step_ms %.3f kernel_ms %s frac %s e2e %.4e (%s ms, h2d %s d2h %s) exact %s numa %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms"), r.get("frac"), e.get("value", 0), e.get("ms_per_step"), e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"), d.get("exact_edge_scores"), d.get("numa")))
except Exception as ex: print("$f: failed", ex)
PY
done
tail -3 gpurun_out/${T}_bench1.err gpurun_out/${T}_bench2.err gpurun_out/${T}_ref.err
