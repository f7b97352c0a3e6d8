#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the profiles/ notes quote.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [regex]"""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|"
                 r"launch__(registers_per_thread|grid_size|block_size|occupancy_limit|shared_mem_per_block_dynamic)|"
                 r"sm__throughput.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__inst_executed.sum$|"
                 r"sm__inst_executed_pipe_(lsu|alu|fma|fmaheavy|xu|uniform|adu|cbu).sum$|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|"
                 r"sm__cycles_elapsed.avg$|smsp__issue_active.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|"
                 r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?.sum$|issue_stalled_.*_per_issue_active.ratio|"
                 r"smsp__inst_executed_op_(shared|global)_(ld|st|atom).sum$|sm__sass_inst_executed_op_(shared|global)|lts__t_bytes.sum$|"
                 r"l1tex__t_bytes.sum$|smsp__cycles_active.avg$|sm__warps_active.avg.per_cycle_active|smsp__thread_inst_executed_per_inst_executed.ratio")
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if pat.search(h):
            print("  %-90s %-14s %s" % (h, units[i], r[i]))
