#!/bin/bash
# L1 pipe breakdown of hc_score_kernel for every variant library (one launch each, 2 M pairs / 25 M candidates):
# tools/ab_ncu.sh  -> gpurun_out/ab_ncu_<variant>.csv
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_global_ld.sum,sm__cycles_elapsed.avg,l1tex__lsu_writeback_active.sum,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,smsp__inst_executed_op_shared_st.sum
mkdir -p gpurun_out
for so in haploconduct_b200/lib/variants/libhc_b200_*.so; do
  n=$(basename $so .so)
  HC_B200_LIB=$PWD/$so timeout 600 ncu --metrics $M --clock-control none -k regex:hc_score_kernel -s 3 -c 1 --csv --log-file gpurun_out/ab_ncu_$n.csv \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --pairs 2000000 --cands 25000000 > /dev/null 2>&1
  echo "== $n"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ab_ncu_$n.csv")) if len(r)>5]
h=rows[0]; i=h.index("Metric Name"); v=h.index("Metric Value")
for r in rows[1:]: print("  %-75s %s"%(r[i], r[v]))
PY
done
