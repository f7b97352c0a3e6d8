#!/bin/bash
# A/B timing of kernel variants on the GPU box: tools/ab_bench.sh [bench args]   (add --e2e to time the host pipeline too)
E2E="--no-e2e"; ARGS=()
for a in "$@"; do if [ "$a" == "--e2e" ]; then E2E=""; else ARGS+=("$a"); fi; done
for so in haploconduct_b200/lib/variants/libhc_b200_*.so; do
  HC_B200_LIB=$PWD/$so timeout 300 python bench.py --no-cpu $E2E --steps 5 "${ARGS[@]}" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d.get('e2e') or {}; print('%-40s value %.4e  kernel_ms %.3f  frac %.4f  step_ms %.3f  e2e_ms %s' % ('$(basename $so)', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'], e.get('ms_per_step')))"
done
