#!/usr/bin/env python
"""Executed warp-instructions and stall samples of a kernel per SOURCE LINE:
   tools/ncu_lines.py <source.csv from `ncu --page source --csv`> <nvdisasm -g -c listing of the same cubin> <mangled-name substring> [top N]"""
import csv, re, sys
from collections import defaultdict
src_csv, dis, fn = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# instruction index -> line, from the listing of the function's .text section
lines, cur, on = [], "?", False
for l in open(dis):
    if l.startswith("\t.section") or l.lstrip().startswith(".section"):
        on = (".text." in l) and (fn in l)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).split("/")[-1]
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = "%s:%s" % (f, m.group(2)) + (" <- %s:%s" % (inl.group(1).split("/")[-1], inl.group(2)) if inl else "")
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
print("listing instructions %d, profile instructions %d" % (len(lines), len(body)))
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0.0])
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
for k, r in enumerate(body):
    key = lines[k] if k < len(lines) else "?"
    a = agg[key]
    a[0] += f(r, "Instructions Executed"); a[1] += f(r, "# Samples"); a[2] += f(r, "stall_long_sb"); a[3] += f(r, "stall_no_inst"); a[4] += f(r, "stall_short_sb")
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("by executed instructions:")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %5.2f%% inst  %5.2f%% samp  long_sb %6d no_inst %6d short_sb %6d | %s" % (100 * a[0] / ti, 100 * a[1] / ts, a[2], a[3], a[4], key))
print("by stall samples:")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("  %5.2f%% inst  %5.2f%% samp  long_sb %6d no_inst %6d short_sb %6d | %s" % (100 * a[0] / ti, 100 * a[1] / ts, a[2], a[3], a[4], key))
