#!/bin/bash
# A/B timing of kernel variants on the GPU box: tools/ab_bench.sh [bench args]
for so in haploconduct_b200/lib/variants/libhc_b200_*.so; do
  HC_B200_LIB=$PWD/$so timeout 300 python bench.py --no-cpu --no-e2e --steps 5 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-40s value %.4e  kernel_ms %.3f  frac %.4f' % ('$(basename $so)', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
