#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu source page: usage
   ncu -i X.ncu-rep --page source --csv | python tools/ncu_source.py [top N]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
tot_wf = sum(f(r, "L1 Wavefronts Shared") for r in body)
print("total warp-instructions %.3e  samples %d  shared wavefronts %.3e" % (tot_inst, tot_samp, tot_wf))
# opcode histogram
from collections import defaultdict
op = defaultdict(float); ops = defaultdict(float)
for r in body:
    s = r[ix["Source"]].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    o = s.split()[0].split(".")[0]
    op[o] += f(r, "Instructions Executed"); ops[o] += f(r, "# Samples")
print("opcode share of executed warp-instructions / of stall samples:")
for o, v in sorted(op.items(), key=lambda kv: -kv[1])[:24]:
    print("  %-10s %6.2f%%  %6.2f%%" % (o, 100 * v / tot_inst, 100 * ops[o] / max(tot_samp, 1)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print("top %d instructions by stall samples:" % n)
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:n]:
    print("  %5.2f%% samp  inst %.2e  wfS %.2e (ideal %.2e)  long_sb %5d short_sb %5d mio %4d wait %4d math %4d | %s" % (
        100 * f(r, "# Samples") / tot_samp, f(r, "Instructions Executed"), f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal"),
        f(r, "stall_long_sb"), f(r, "stall_short_sb"), f(r, "stall_mio"), f(r, "stall_wait"), f(r, "stall_math"), r[ix["Source"]].strip()[:70]))
lds = [r for r in body if r[ix["Source"]].strip().split()[-0:1] and "LDS" in r[ix["Source"]]]
wf = sum(f(r, "L1 Wavefronts Shared") for r in lds); idl = sum(f(r, "L1 Wavefronts Shared Ideal") for r in lds); ie = sum(f(r, "Instructions Executed") for r in lds)
print("LDS: %.3e instr, %.3e wavefronts (%.2f per instr), ideal %.3e" % (ie, wf, wf / max(ie, 1), idl))
